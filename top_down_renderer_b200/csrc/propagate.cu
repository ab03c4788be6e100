// propagate.cu — SURVEY section 8f rank 1: the motion model on the device-resident particle set.
//
// Reference: StateParticle::propagate (src/state_particle.cpp:57-78), ParticleFilter::propagate
// (src/particle_filter.cpp:86-92).  Per particle, in this order:
//     (gx, gy) = Rotation2D(theta) * trans ;  dx += gx ; dy += gy ;  dist = |(gx, gy)|
//     theta += N(0, theta_cov*dist) + omega ;  dx += N(0, pos_cov*dist) ;  dy += N(0, pos_cov*dist)
//     if (!scale_freeze) scale *= N(1, min(2/dist, 0.02)) ;  last_dist = |old (dx, dy) - new (dx, dy)|
// The reference draws from ONE shared std::mt19937 in particle order (3-4 normal_distribution<float> calls per
// particle, each distribution object fresh): a sequential stream that cannot be reproduced in parallel, and with 1e6
// device-resident particles the host loop costs a 56 MB round trip plus tens of ms of RNG per scan.  Two entry points:
//   * injected variates (tdr_pf_propagate): the caller supplies the STANDARD normal variates z (4 per particle: theta,
//     dx, dy, scale — what libstdc++'s normal_distribution returns before `* stddev + mean`), e.g. drawn with the
//     reference's own RNG calls; the kernel applies them with the reference's fp32 operation order.  This is the
//     parity path: identical z in, identical states out (cos/sin: double-precision evaluation rounded to fp32,
//     which is what glibc's cosf/sinf return up to rare 1-ulp double-rounding cases; tolerance in the tests).
//   * device RNG (tdr_pf_propagate_rng): Philox-4x32-10 keyed by (seed, step), counter = particle index, Box-Muller
//     -> the same 4 variates without any host traffic.  Same distribution, different stream (statistical tests).
#include "tdr_ctx.cuh"
#include "tdr_math.cuh"

namespace tdr {

struct PropParams {
  float *dx, *dy, *theta, *scale, *last_dist;
  long long n;
  float tx, ty, omega, pos_cov, theta_cov;
  int scale_freeze;
  const float* z;                 // injected: [n][4]
  unsigned long long seed, step;  // device RNG
  float* z_out;                   // optional: the variates used, [n][4]
};

// Philox-4x32-10 (Salmon et al., SC'11)
__device__ __forceinline__ void philox4x32(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t lo0 = 0xD2511F53u * c[0], hi0 = __umulhi(0xD2511F53u, c[0]);
    const uint32_t lo1 = 0xCD9E8D57u * c[2], hi1 = __umulhi(0xCD9E8D57u, c[2]);
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
// two standard normals from two 32-bit words (Box-Muller in double: u1 in (0, 1], u2 in [0, 1))
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float* z0, float* z1) {
  const double u1 = ((double)a + 1.0) * (1.0 / 4294967296.0), u2 = (double)b * (1.0 / 4294967296.0);
  const double r = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  *z0 = (float)(r * c); *z1 = (float)(r * s);
}

template <bool RNG>
__global__ void k_propagate(PropParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  float z[4];
  if (RNG) {
    uint32_t c[4] = {(uint32_t)i, (uint32_t)((unsigned long long)i >> 32), (uint32_t)p.step, (uint32_t)(p.step >> 32)};
    philox4x32(c, (uint32_t)p.seed, (uint32_t)(p.seed >> 32));
    box_muller(c[0], c[1], &z[0], &z[1]);
    box_muller(c[2], c[3], &z[2], &z[3]);
  } else {
    const float4 q = reinterpret_cast<const float4*>(p.z)[i];
    z[0] = q.x; z[1] = q.y; z[2] = q.z; z[3] = q.w;
  }
  if (p.z_out) reinterpret_cast<float4*>(p.z_out)[i] = make_float4(z[0], z[1], z[2], z[3]);
  const float th = p.theta[i];
  const float c = (float)cos((double)th), s = (float)sin((double)th);                 // std::cos / std::sin (float) :58
  const float gx = TDR_FSUB(TDR_FMUL(c, p.tx), TDR_FMUL(s, p.ty));                     // Rotation2D(theta) * trans
  const float gy = TDR_FADD(TDR_FMUL(s, p.tx), TDR_FMUL(c, p.ty));
  const float lx = p.dx[i], ly = p.dy[i];                                              // last_pos :59
  float dx = TDR_FADD(lx, gx), dy = TDR_FADD(ly, gy);                                  // :60-61
  const float dist = TDR_FSQRT(TDR_FADD(TDR_FMUL(gx, gx), TDR_FMUL(gy, gy)));          // :63
  const float sd_pos = TDR_FMUL(p.pos_cov, dist), sd_th = TDR_FMUL(p.theta_cov, dist); // :64-65
  // normal_distribution returns z * stddev + mean
  p.theta[i] = TDR_FADD(th, TDR_FADD(TDR_FADD(TDR_FMUL(z[0], sd_th), 0.f), p.omega));  // :67
  dx = TDR_FADD(dx, TDR_FADD(TDR_FMUL(z[1], sd_pos), 0.f));                            // :68
  dy = TDR_FADD(dy, TDR_FADD(TDR_FMUL(z[2], sd_pos), 0.f));                            // :69
  p.dx[i] = dx; p.dy[i] = dy;
  if (!p.scale_freeze) {                                                               // :71-74
    const double m = 2.0 / (double)dist;
    const float sd_sc = (float)(0.02 < m ? 0.02 : m);                                  // std::min(2./dist, 0.02)
    p.scale[i] = TDR_FMUL(p.scale[i], TDR_FADD(TDR_FMUL(z[3], sd_sc), 1.f));
  }
  const float mx = TDR_FSUB(lx, dx), my = TDR_FSUB(ly, dy);                            // :76-77
  p.last_dist[i] = TDR_FSQRT(TDR_FADD(TDR_FMUL(mx, mx), TDR_FMUL(my, my)));
}

// z_src: device pointer to [n][4] standard normal variates (rng == false); z_out_dev optional
int propagate(tdr_ctx* ctx, float tx, float ty, float omega, int scale_freeze, float pos_cov, float theta_cov, const float* z_dev,
              bool rng, unsigned long long seed, unsigned long long step, float* z_out_dev) {
  Particles& pt = ctx->part[ctx->cur];
  TDR_REQUIRE(pt.n > 0, TDR_ESTATE, "no particles");
  PropParams p; memset(&p, 0, sizeof(p));
  p.dx = pt.dx.as<float>(); p.dy = pt.dy.as<float>(); p.theta = pt.theta.as<float>(); p.scale = pt.scale.as<float>();
  p.last_dist = pt.last_dist.as<float>(); p.n = pt.n;
  p.tx = tx; p.ty = ty; p.omega = omega; p.pos_cov = pos_cov; p.theta_cov = theta_cov;
  p.scale_freeze = scale_freeze; p.z = z_dev; p.seed = seed; p.step = step; p.z_out = z_out_dev;
  const unsigned blocks = (unsigned)((pt.n + 255) / 256);
  if (rng) k_propagate<true><<<blocks, 256, 0, ctx->stream>>>(p);
  else k_propagate<false><<<blocks, 256, 0, ctx->stream>>>(p);
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  ctx->have_argmax = false;
  return TDR_OK;
}

}  // namespace tdr
