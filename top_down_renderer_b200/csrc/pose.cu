// pose.cu — a13: mean / covariance / max-likelihood pose, and the grid arg-min.
//
// Reference: StateParticle::mlState (src/state_particle.cpp:98-102), ParticleFilter::meanLikelihood
// (src/particle_filter.cpp:191-203), computeMeanCov (:205-220), maxLikelihood (:222-224),
// computeCov (:226-236).
//
// The x / y / scale means are sequential fp32 sums in the reference; at N >= 1e4 their rounding is
// larger than the 1 mm parity bar, so they are reproduced order-exactly (k_exact_seq).  The circular
// mean of theta and the covariance are accumulated in double (inside the 0.01 deg / 1e-4 bars for
// any summation order; DESIGN.md "pose").
#include "tdr_ctx.cuh"
#include "tdr_math.cuh"

namespace tdr {

struct PartPtrs { const float *ix, *iy, *dx, *dy, *th, *sc; };

__device__ __forceinline__ void ml_state(const PartPtrs& p, long long i, float s[4]) {
  float sc = p.sc[i];
  s[0] = TDR_FADD(TDR_FMUL(p.dx[i], sc), p.ix[i]);
  s[1] = TDR_FADD(TDR_FMUL(p.dy[i], sc), p.iy[i]);
  s[2] = p.th[i];
  s[3] = sc;
}

// columns: mx | my | ms (each n floats), plus double sums of cosf / sinf
__global__ void k_ml_columns(PartPtrs p, long long n, float* __restrict__ cols, double* __restrict__ trig) {
  double cs = 0.0, sn = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float s[4]; ml_state(p, i, s);
    cols[i] = s[0]; cols[n + i] = s[1]; cols[2 * n + i] = s[3];
    cs += (double)cosf(s[2]); sn += (double)sinf(s[2]);
  }
  for (int o = 16; o > 0; o >>= 1) { cs += __shfl_xor_sync(0xffffffffu, cs, o); sn += __shfl_xor_sync(0xffffffffu, sn, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(trig, cs); atomicAdd(trig + 1, sn); }
}

// scal[SC_POSE + 0..3] = mean state ; totals in scal[SC_POSE + 8..10]
__global__ void k_pose_mean(float* scal, const double* trig, long long n) {
  const float fn = (float)(unsigned long long)n;
  float* o = scal + SC_POSE;
  o[0] = TDR_FDIV(o[8], fn); o[1] = TDR_FDIV(o[9], fn); o[3] = TDR_FDIV(o[10], fn);   // :201
  float cs = (float)trig[0], sn = (float)trig[1];
  o[2] = fdlibm_atan2f(TDR_FDIV(sn, fn), TDR_FDIV(cs, fn));                            // :202
}

// scal[SC_POSE + 4..7] = ml state of particle argmax in buffer `mlp`
__global__ void k_pose_ml(float* scal, PartPtrs mlp, long long n_ml) {
  int a = reinterpret_cast<const int*>(scal)[SC_ARGMAX];
  float s[4] = {0, 0, 0, 0};
  if (a >= 0 && a < n_ml) ml_state(mlp, a, s);
  for (int k = 0; k < 4; k++) scal[SC_POSE + 4 + k] = s[k];
}

// 10 unique entries of sum delta delta^T about ref = scal[ref_slot..+3]; acc: 10 doubles
__global__ void k_cov_accum(PartPtrs p, long long n, const float* __restrict__ scal, int ref_slot, double* __restrict__ acc) {
  const float r0 = scal[ref_slot], r1 = scal[ref_slot + 1], r2 = scal[ref_slot + 2], r3 = scal[ref_slot + 3];
  double a[10];
#pragma unroll
  for (int k = 0; k < 10; k++) a[k] = 0.0;
  const double PI = 3.14159265358979323846;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float s[4]; ml_state(p, i, s);
    s[0] = TDR_FSUB(s[0], r0); s[1] = TDR_FSUB(s[1], r1); s[2] = TDR_FSUB(s[2], r2); s[3] = TDR_FSUB(s[3], r3);
    int guard = 0;
    while ((double)s[2] > PI && guard++ < 64) s[2] = (float)((double)s[2] - 2 * PI);     // :215-216
    while ((double)s[2] < -PI && guard++ < 64) s[2] = (float)((double)s[2] + 2 * PI);
    int k = 0;
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
      for (int r = 0; r <= c; r++) a[k++] += (double)TDR_FMUL(s[r], s[c]);
  }
#pragma unroll
  for (int k = 0; k < 10; k++) {
    double v = a[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(acc + k, v);
  }
}

// cov (col-major 4x4) = acc / (n-1) into scal[out_slot..+15]
__global__ void k_cov_finish(const double* acc, long long n, float* scal, int out_slot) {
  const float den = (float)(unsigned long long)(n - 1);    // cov /= particles_.size()-1  (:219)
  int k = 0;
  for (int c = 0; c < 4; c++)
    for (int r = 0; r <= c; r++) {
      float v = TDR_FDIV((float)acc[k++], den);
      scal[out_slot + c * 4 + r] = v; scal[out_slot + r * 4 + c] = v;
    }
}

static PartPtrs ptrs_of(const Particles& p) {
  PartPtrs q; q.ix = p.init_x.as<float>(); q.iy = p.init_y.as<float>(); q.dx = p.dx.as<float>(); q.dy = p.dy.as<float>();
  q.th = p.theta.as<float>(); q.sc = p.scale.as<float>();
  return q;
}

// max_likelihood_particle_->mlState() (particle_filter.cpp:145-147,222-224): cached when the arg-max is
// taken, because the particle it points at belongs to the pre-resample set.
int cache_ml_state(tdr_ctx* ctx, const Particles& src) {
  k_pose_ml<<<1, 1, 0, ctx->stream>>>(ctx->scal.as<float>(), ptrs_of(src), src.n);
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  ctx->have_argmax = true;
  return TDR_OK;
}

int pose_of(tdr_ctx* ctx, Particles& pt, float* mean, float* cov_mean, float* ml, float* cov_ml) {
  const long long n = pt.n;
  TDR_REQUIRE(n > 0, TDR_ESTATE, "no particles");
  if (int e = ctx->scal.reserve(SC_TOTAL * 4)) return e;
  if (int e = ctx->pose_tmp.reserve((size_t)n * 3 * 4 + 64 * 8)) return e;
  float* scal = ctx->scal.as<float>();
  float* cols = ctx->pose_tmp.as<float>();
  double* dacc = reinterpret_cast<double*>(scal + SC_POSE + 64);   // 32 doubles: trig[2] | cov_mean[10] | cov_ml[10]
  TDR_CUDA(cudaMemsetAsync(dacc, 0, 32 * 8, ctx->stream));
  PartPtrs p = ptrs_of(pt);
  const int blocks = (int)((n + 255) / 256 < ctx->sm_count * 8 ? (n + 255) / 256 : ctx->sm_count * 8);
  k_ml_columns<<<blocks, 256, 0, ctx->stream>>>(p, n, cols, dacc);
  count_launch(ctx);
  const float* cptr[3] = {cols, cols + n, cols + 2 * n};
  if (int e = exact_sums(ctx, cptr, n, 3, scal + SC_POSE + 8)) return e;
  k_pose_mean<<<1, 1, 0, ctx->stream>>>(scal, dacc, n);
  count_launch(ctx);
  if (cov_mean) {
    k_cov_accum<<<blocks, 256, 0, ctx->stream>>>(p, n, scal, SC_POSE + 0, dacc + 2);
    k_cov_finish<<<1, 1, 0, ctx->stream>>>(dacc + 2, n, scal, SC_POSE + 16);
    count_launch(ctx, 2);
  }
  if (ml || cov_ml) {
    TDR_REQUIRE(ctx->have_argmax, TDR_ESTATE, "max-likelihood pose needs a previous tdr_pf_normalize");
    if (cov_ml) {
      k_cov_accum<<<blocks, 256, 0, ctx->stream>>>(p, n, scal, SC_POSE + 4, dacc + 12);
      k_cov_finish<<<1, 1, 0, ctx->stream>>>(dacc + 12, n, scal, SC_POSE + 32);
      count_launch(ctx, 2);
    }
  }
  TDR_CUDA(cudaGetLastError());
  float host[48];
  TDR_CUDA(cudaMemcpyAsync(host, scal + SC_POSE, 48 * 4, cudaMemcpyDeviceToHost, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  if (mean) for (int k = 0; k < 4; k++) mean[k] = host[k];
  if (ml) for (int k = 0; k < 4; k++) ml[k] = host[4 + k];
  if (cov_mean) for (int k = 0; k < 16; k++) cov_mean[k] = host[16 + k];
  if (cov_ml) for (int k = 0; k < 16; k++) cov_ml[k] = host[32 + k];
  return TDR_OK;
}

// ---- the strided sample matrix of ParticleFilter::computeGMM (particle_filter.cpp:262-272):
// sample i = mlState of particle min(n - 1, i * n / num_samples) as (x, y, 50 cos theta, 50 sin theta) in double
__global__ void k_gmm_samples(PartPtrs p, long long n, int num_samples, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num_samples) return;
  long long idx = (long long)i * n / num_samples;
  if (idx > n - 1) idx = n - 1;
  float s[4]; ml_state(p, idx, s);
  out[4 * i + 0] = (double)s[0];
  out[4 * i + 1] = (double)s[1];
  out[4 * i + 2] = (double)TDR_FMUL(50.f, (float)cos((double)s[2]));           // 50 * cos(float) is a float product
  out[4 * i + 3] = (double)TDR_FMUL(50.f, (float)sin((double)s[2]));
}
int gmm_samples(tdr_ctx* ctx, int num_samples, double* samples_host) {
  Particles& pt = ctx->part[ctx->cur];
  TDR_REQUIRE(pt.n > 0 && num_samples > 0 && samples_host, TDR_EINVAL, "bad sample request");
  if (int e = ctx->scratch.reserve((size_t)num_samples * 32)) return e;
  PartPtrs p{pt.init_x.as<float>(), pt.init_y.as<float>(), pt.dx.as<float>(), pt.dy.as<float>(), pt.theta.as<float>(), pt.scale.as<float>()};
  k_gmm_samples<<<(num_samples + 127) / 128, 128, 0, ctx->stream>>>(p, pt.n, num_samples, ctx->scratch.as<double>());
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  TDR_CUDA(cudaMemcpyAsync(samples_host, ctx->scratch.p, (size_t)num_samples * 32, cudaMemcpyDeviceToHost, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  return TDR_OK;
}

// ---- exhaustive grid: (min cost, first index) over n*n_shifts costs, NaN never wins
__global__ void k_grid_best(const float* __restrict__ costs, long long n, unsigned long long* __restrict__ best) {
  unsigned long long loc = ~0ull;   // key = (ordered bits << 32 | low index bits)... index may exceed 32 bits -> two-stage
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = costs[i];
    if (v == v) {
      uint32_t u = __float_as_uint(v);
      u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
      unsigned long long key = ((unsigned long long)u << 32) | (unsigned long long)(uint32_t)(i & 0xffffffffll);
      if (key < loc) loc = key;
    }
  }
  for (int o = 16; o > 0; o >>= 1) { unsigned long long t = __shfl_xor_sync(0xffffffffu, loc, o); if (t < loc) loc = t; }
  if ((threadIdx.x & 31) == 0) atomicMin(best, loc);
}

int grid_best(tdr_ctx* ctx, const float* dev_costs, long long n, float* best_cost, long long* best_index) {
  TDR_REQUIRE(n > 0 && dev_costs, TDR_ESTATE, "no grid costs");
  TDR_REQUIRE(n < (1ll << 32), TDR_EUNSUPPORTED, "grid too large for tdr_grid_best");
  if (int e = ctx->scal.reserve(SC_TOTAL * 4)) return e;
  unsigned long long* key = reinterpret_cast<unsigned long long*>(ctx->scal.as<float>() + SC_DBL) + 6;
  TDR_CUDA(cudaMemsetAsync(key, 0xff, 8, ctx->stream));
  const int blocks = (int)((n + 255) / 256 < ctx->sm_count * 8 ? (n + 255) / 256 : ctx->sm_count * 8);
  k_grid_best<<<blocks, 256, 0, ctx->stream>>>(dev_costs, n, key);
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  unsigned long long h = 0;
  TDR_CUDA(cudaMemcpyAsync(&h, key, 8, cudaMemcpyDeviceToHost, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h == ~0ull) { if (best_cost) *best_cost = NAN; if (best_index) *best_index = -1; return TDR_OK; }
  uint32_t u = (uint32_t)(h >> 32);
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  float v; memcpy(&v, &u, 4);
  if (best_cost) *best_cost = v;
  if (best_index) *best_index = (long long)(h & 0xffffffffull);
  return TDR_OK;
}

}  // namespace tdr
