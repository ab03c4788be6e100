// score_mma_list.cu — a7/a9/a10 as a tcgen05 gather-GEMM for a SHORT candidate list (the 40-shift theta search).
//
// Reference: TopDownMapPolar::getLocalMap (src/top_down_map_polar.cpp:21-53), StateParticle::getCostForRot
// (src/state_particle.cpp:112-155), StateParticle::computeWeight (:157-219).
//
// Same gather and operand format as score_mma.cu (one 32-byte fp16 hi/lo map record per lattice cell = one K = 16
// MMA step), but the scan operand holds ONLY the candidate shifts: per cell a [2*S_pad x 16] block
//     rows n < S      : class counts at the shifted angle (hi and lo slots)
//     row  n == S     : 1 at slot 7 -> D = number of known cells
//     rows S_pad + s  : class-summed count at slot 7 -> the normalisation
// precomputed per scan and streamed per stage with cp.async.bulk.  N = 96 accumulator columns for up to 47
// candidates instead of the 224 the all-shifts kernel needs, so two CTAs (two independent gather/MMA pipelines)
// share an SM: 12.5 ms against 14.2 ms for 1e6 hypotheses x 40 shifts on one B200.  The price is the operand
// stream (7.7 MB per 128 x T hypotheses out of L2), which is why many-shift searches use score_mma.cu instead.
#include "mma_common.cuh"

namespace tdr {

// scan operand, per cell p: [kc = 2][n = N][8 halfs]  (K-major canonical layout, LBO = N*16 B, SBO = 128 B)
// one thread per (p, n).  maxcount: device int, max class-summed count seen (fp16 integers are exact up to 2048).
static __global__ void k_build_scan_operand(const float* __restrict__ img, int C, int n_theta, int n_r, int P, int P_pad,
                                     const int32_t* __restrict__ shifts, int S, int S_pad, int N,
                                     uint4* __restrict__ out, int* __restrict__ maxcount) {
  long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (long long)P_pad * N) return;
  const int p = (int)(id / N), n = (int)(id - (long long)p * N);
  unsigned short h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  bool lo_copy = false;
  if (p < P) {
    const int r = p / n_theta, th = p - r * n_theta;
    if (n < S || (n >= S_pad && n < S_pad + S)) {
      const int s = n < S ? n : n - S_pad;
      int t2 = th + shifts[s];
      t2 %= n_theta; if (t2 < 0) t2 += n_theta;
      const int cell = r * n_theta + t2;          // scan row (theta + shift) pairs with map row theta
      float tot = 0.f;
      for (int c = 0; c < C; c++) {
        float v = img[(size_t)c * P + cell];
        tot += v;
        if (n < S) h[c] = __half_as_ushort(__float2half_rn(v));
      }
      if (n < S) lo_copy = true;
      else h[7] = __half_as_ushort(__float2half_rn(tot));
      if (n == S_pad) atomicMax(maxcount, (int)tot);     // shift[0] is a bijection of the cells: global max of tot
    } else if (n == S) {
      h[7] = 0x3C00;                              // 1.0: counts the known cells
    }
  }
  uint4 q;
  q.x = h[0] | ((uint32_t)h[1] << 16); q.y = h[2] | ((uint32_t)h[3] << 16);
  q.z = h[4] | ((uint32_t)h[5] << 16); q.w = h[6] | ((uint32_t)h[7] << 16);
  uint4 z = make_uint4(0, 0, 0, 0);
  uint4* cellbase = out + (size_t)p * N * 2;      // N*32 B per cell = 2N uint4
  cellbase[n] = q;                                // kc = 0: hi slots
  cellbase[N + n] = lo_copy ? q : z;              // kc = 1: lo slots see the same counts (slot 7 lo is never set)
}


struct ListParams {
  const uint4* map16; GatherGeom geom; float resolution; int tab_scaled;
  const float2* tab; int P, P_pad; float res;
  const uint4* bop;
  const int* perm; long long n_work;
  // particle mode
  const float *init_x, *init_y, *dx, *dy; float* theta; const float* scale; uint8_t* have_init; float* weights;
  int force_on_map; float map_w, map_h; int scale_gate; double scale_lo, scale_hi; float regularization;
  const float* thetas; int n_shifts;
  // grid mode
  const float* centers; float grid_scale; float* costs;
  // device-side precondition: fp16 holds integer counts exactly up to 2048.  Above that the kernel leaves at once and
  // the guarded CUDA-core launch that follows does the work (no host round trip to decide)
  const int* maxcount; int* bailed;
};


// T = 128-hypothesis tiles per CTA; R = gather threads per hypothesis row (the R threads of a row take turns
// stage by stage, so the loads in flight per SM double without doubling the hypotheses — and their map
// footprint — that are in flight together)
// ATM = the gathered records go to TENSOR MEMORY instead of shared memory (tcgen05.st, 8 columns = 16 fp16 per row
// and cell) and the MMA takes its A operand from there: the copy into the operand no longer costs L1 data-pipe
// wavefronts (a quarter of the pipe that bounds this kernel), TMEM writes run at 256 B/clk beside it.
template <int N, int T, int R, bool ATM> struct ListCfg {
  static const int kThreads = 128 * T * R + 64;
  static const int kTmemCols = ATM ? 512 : (T * N <= 128 ? 128 : (T * N <= 256 ? 256 : 512));
  static const int kByTmem = 512 / kTmemCols, kByRegs = 65536 / (kThreads * 88) < 1 ? 1 : 65536 / (kThreads * 88);
  static const int kCtasPerSm = kByTmem < kByRegs ? kByTmem : kByRegs;
  static const int kABytes = ATM ? 0 : MMA_G * T * A_TILE;   // per stage
  static const int kBBytes = MMA_G * N * 32;              // per stage
  static const int kStageBytes = kABytes + kBBytes;
  static const int kBudget = (216 * 1024) / kCtasPerSm - 1536;
  static const int kACols = MMA_G * T * 8;                // TMEM columns of one stage of A (ATM)
  static const int kStagesTm = (512 - T * N) / kACols > 16 ? 16 : (512 - T * N) / kACols;
  static const int kStagesSm = kBudget / kStageBytes > 8 ? 8 : kBudget / kStageBytes;
  static const int kStages = ATM ? kStagesTm : kStagesSm;
  static const int kSmem = kStages * kStageBytes + 512;     // + barriers (2 * kStages + 1) and the TMEM base word
};

template <int N, int T, int R, bool ATM>
__global__ void __launch_bounds__(128 * T * R + 64, ListCfg<N, T, R, ATM>::kCtasPerSm) k_score_mma_list(ListParams sp) {
  using Cfg = ListCfg<N, T, R, ATM>;
  constexpr int GW = 4 * T * R;        // gather warps
  constexpr int NS = Cfg::kStages;
  constexpr int S_PAD = N / 2;
  if (*sp.maxcount > MMA_MAX_EXACT_COUNT) {          // grid-uniform, before anything is allocated
    if (blockIdx.x == 0 && threadIdx.x == 0) *sp.bailed = 1;
    return;
  }
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sA = smem;                                  // [NS][G][T] tiles of A_TILE bytes
  unsigned char* sB = smem + (size_t)NS * Cfg::kABytes;       // [NS][G][kc 2][N][16 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)NS * Cfg::kStageBytes);   // full[NS] empty[NS] accum
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * NS + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + NS), bar_accum = smem_u32(bars + 2 * NS);

  if (warp == GW + 1) tmem_alloc(smem_u32(s_tmem), Cfg::kTmemCols);
  if (tid == 0) {
    for (int s = 0; s < NS; s++) { mbar_init(bar_full + 8 * s, 4 * T + 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_accum, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  const long long per_batch = 128 * T;
  const long long n_batches = (sp.n_work + per_batch - 1) / per_batch;
  const int K_ITERS = sp.P_pad / MMA_G;
  uint32_t it = 0;                 // pipeline iteration counter, continues across batches (same sequence in every role)
  uint32_t local_batch = 0;

  if (warp < GW) {
    // =========================== gather + epilogue ===========================
    const int sub = warp / (4 * T);                      // which of the R threads of a row this is
    const int t = (warp % (4 * T)) >> 2, m = tid & 127;
    const unsigned char* map_bytes = reinterpret_cast<const unsigned char*>(sp.map16);
    for (long long batch = blockIdx.x; batch < n_batches; batch += gridDim.x, local_batch++) {
      const long long slot = batch * per_batch + (tid % (128 * T));
      long long i = -1;
      if (slot < sp.n_work) i = sp.perm ? (long long)sp.perm[slot] : slot;
      float cx = 0.f, cy = 0.f, sc = 1.f;
      bool active = false, gated = false;
      if (i >= 0) {
        if (sp.centers) { cx = sp.centers[2 * i]; cy = sp.centers[2 * i + 1]; sc = sp.grid_scale; active = true; }
        else {
          sc = sp.scale[i];
          cx = TDR_FADD(TDR_FMUL(sp.dx[i], sc), sp.init_x[i]);
          cy = TDR_FADD(TDR_FMUL(sp.dy[i], sc), sp.init_y[i]);
          if (sp.force_on_map && (cx < 0.f || cy < 0.f || cx > sp.map_w || cy > sp.map_h)) gated = true;      // :163-168
          if (sp.scale_gate && ((double)sc < sp.scale_lo || (double)sc > sp.scale_hi)) gated = true;          // :169-176
          active = !gated;
        }
      }
      const float oy = TDR_FDIV(cy, sp.resolution), ox = TDR_FDIV(cx, sp.resolution);

      // Each thread pulls the whole 32-byte record of ITS hypothesis with one 256-bit load (one sector, one L1
      // wavefront; measured 0.95 records/clk/SM from L2 against 0.42 for 2 x LDG.128 — tools/gather_bench.cu).
      auto load_stage = [&](int k, uint4 (&rec)[MMA_G][2]) {
#pragma unroll
        for (int g = 0; g < MMA_G; g++) {
          const int p = k * MMA_G + g;
          uint32_t off = sp.geom.zero_rec;
          if (p < sp.P) {                                  // warp-uniform (the last stages are padding)
            const float2 tb = c_tab[p];
            float vy, vx;
            if (sp.tab_scaled) { vy = TDR_FADD(tb.x, oy); vx = TDR_FADD(tb.y, ox); }          // grid: the table holds (tab * scale) * res
            else { vy = TDR_FADD(TDR_FMUL(TDR_FMUL(tb.x, sc), sp.res), oy); vx = TDR_FADD(TDR_FMUL(TDR_FMUL(tb.y, sc), sp.res), ox); }
            off = lattice_record(vy, vx, sp.geom);
            if (!active) off = sp.geom.zero_rec;
          }
          ldg256(map_bytes + (size_t)off * 32, rec[g][0], rec[g][1]);
        }
      };
      auto store_stage = [&](uint32_t iter, const uint4 (&rec)[MMA_G][2]) {
        const uint32_t st = iter % NS, ph = (iter / NS) & 1u;
        mbar_wait(bar_empty + 8 * st, ph ^ 1u);
        if (ATM) {
          // row m of the A tile = TMEM lane m (this warp's quarter), 8 columns = the record's 16 fp16 in K order
          const uint32_t tcol = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(T * N) +
                                (uint32_t)(st * Cfg::kACols + t * 8);
#pragma unroll
          for (int g = 0; g < MMA_G; g++) tmem_st8(tcol + (uint32_t)(g * T * 8), rec[g][0], rec[g][1]);
          tmem_wait_st();
          tc_fence_before();
        } else {
          unsigned char* base = sA + (size_t)st * Cfg::kABytes + (size_t)t * A_TILE + (size_t)m * 16;
#pragma unroll
          for (int g = 0; g < MMA_G; g++) {
            *reinterpret_cast<uint4*>(base + (size_t)g * T * A_TILE) = rec[g][0];            // K chunk 0: hi halves
            *reinterpret_cast<uint4*>(base + (size_t)g * T * A_TILE + A_LBO) = rec[g][1];    // K chunk 1: lo halves
          }
          fence_proxy_async();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full + 8 * st);
      };

      uint4 ra[MMA_G][2], rb[MMA_G][2];
      if (sub < K_ITERS) load_stage(sub, ra);
#pragma unroll 1
      for (int k = sub; k < K_ITERS; k += 2 * R) {       // this thread's stages: sub, sub + R, ... (two in flight)
        if (k + R < K_ITERS) load_stage(k + R, rb);
        store_stage(it + k, ra);
        if (k + 2 * R < K_ITERS) load_stage(k + 2 * R, ra);
        if (k + R < K_ITERS) store_stage(it + k + R, rb);
      }
      it += K_ITERS;
      if (sub != 0) continue;                            // the first thread of each row owns the epilogue

      // ---- epilogue: this thread's accumulator row
      mbar_wait(bar_accum, local_batch & 1u);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(t * N);
      uint32_t vc[16], vn[16];
      const uint32_t kraw = tmem_ld1(trow + (uint32_t)sp.n_shifts);
      tmem_wait_ld();
      const float known = __uint_as_float(kraw);
      const bool unknown = (double)TDR_FDIV(known, (float)sp.P) < 0.5;                       // :117-120
      float best = 3.402823466e+38f, best_theta = 0.f;                                       // :193-204
#pragma unroll 1
      for (int ch = 0; ch * 16 < sp.n_shifts; ch++) {
        tmem_ld16(trow + (uint32_t)(ch * 16), vc);
        tmem_ld16(trow + (uint32_t)(S_PAD + ch * 16), vn);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; j++) {
          const int s = ch * 16 + j;
          if (s < sp.n_shifts) {
            float cost = unknown ? __int_as_float(0x7fc00000)
                                 : TDR_FDIV(TDR_FMUL(__uint_as_float(vc[j]), 0.01f), __uint_as_float(vn[j]));   // :137,154
            if (sp.costs && i >= 0) sp.costs[i * sp.n_shifts + s] = cost;
            if (cost < best) { best = cost; best_theta = sp.thetas ? sp.thetas[s] : 0.f; }
          }
        }
      }
      if (i >= 0 && !sp.centers) {
        if (gated) sp.weights[i] = 0.f;
        else {
          sp.theta[i] = best_theta;
          sp.have_init[i] = 1;
          sp.weights[i] = (float)(1.0 / (double)TDR_FADD(best, sp.regularization));             // :212
        }
      }
      tc_fence_before();           // TMEM reads are done before the next batch's first full-barrier arrive
    }
  } else if (warp == GW) {
    // =========================== scan-operand loader ===========================
    // whole warp, warp-uniform control flow; the elected lane issues the copies (see elect_one)
    const bool leader = elect_one();
    const uint32_t sB_u = smem_u32(sB);
    uint32_t st = 0, ph = 0;
    for (long long batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
      const unsigned char* src = reinterpret_cast<const unsigned char*>(sp.bop);
      for (int k = 0; k < K_ITERS; k++, src += Cfg::kBBytes) {
        mbar_wait(bar_empty + 8 * st, ph ^ 1u);
        if (leader) {
          mbar_expect_tx(bar_full + 8 * st, Cfg::kBBytes);
          bulk_g2s(sB_u + st * Cfg::kBBytes, src, Cfg::kBBytes, bar_full + 8 * st);
        }
        __syncwarp();
        if (++st == NS) { st = 0; ph ^= 1u; }
      }
    }
  } else {
    // =========================== MMA issuer ===========================
    // whole warp, warp-uniform control flow, incremental slot / parity counters; the elected lane issues.
    // instruction descriptor: D = f32, A = B = f16, both K-major, N, M = 128
    const bool leader = elect_one();
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
    uint32_t st = 0, ph = 0;
    for (long long batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
      for (int k = 0; k < K_ITERS; k++) {
        mbar_wait(bar_full + 8 * st, ph);
        tc_fence_after();
        if (leader) {
          const uint64_t bdesc0 = umma_desc(sB_u + st * Cfg::kBBytes, N * 16, 128);
#pragma unroll
          for (int g = 0; g < MMA_G; g++) {
            const uint64_t bdesc = bdesc0 + (uint64_t)(g * (N * 32 / 16));
#pragma unroll
            for (int tt = 0; tt < T; tt++) {
              if (ATM) {
                umma_f16_ts(tmem_base + (uint32_t)(tt * N), tmem_base + (uint32_t)(T * N + (g * T + tt) * 8) + st * Cfg::kACols,
                            bdesc, idesc, (k > 0 || g > 0) ? 1u : 0u);
              } else {
                const uint64_t adesc = umma_desc(sA_u + st * Cfg::kABytes + (g * T + tt) * A_TILE, A_LBO, 128);
                umma_f16(tmem_base + (uint32_t)(tt * N), adesc, bdesc, idesc, (k > 0 || g > 0) ? 1u : 0u);
              }
            }
          }
          umma_commit(bar_empty + 8 * st);        // implies tcgen05.fence::before_thread_sync
          if (k == K_ITERS - 1) umma_commit(bar_accum);
        }
        __syncwarp();
        if (++st == NS) { st = 0; ph ^= 1u; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == GW + 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static bool list_usable(tdr_ctx* ctx, int n_shifts) {
  if (ctx->score_impl == 1) return false;
  if (n_shifts < 1 || n_shifts > 47) return false;        // longer lists: the all-shifts kernel (score_mma.cu)
  for (int c = 0; c < ctx->C; c++) {
    float w = ctx->fp.class_weights[c];
    if (!(w >= 0.f) || w * 50.f > 60000.f) return false;     // fp16 range of w_c * dist_c (dist <= 50)
  }
  return true;
}

// returns TDR_OK and sets *used = true when the tensor-core path ran; *used = false -> caller falls back
int score_mma_list(tdr_ctx* ctx, float res, bool grid_mode, long long n_items, float grid_scale, const int32_t* dev_shifts,
              int n_shifts, bool* used) {
  *used = false;
  if (!list_usable(ctx, n_shifts)) return TDR_OK;
  const int P = ctx->n_theta * ctx->n_r;
  if (P > MMA_TAB_MAX) return TDR_OK;
  const int S_pad = n_shifts + 1 <= 48 ? 48 : 112;
  const int N = 2 * S_pad;
  const int P_pad = (P + 2 * MMA_G - 1) / (2 * MMA_G) * (2 * MMA_G);     // even number of stages keeps the 2x unroll simple
  // ---- scan operand (+ max count check: fp16 integers are exact up to 2048)
  if (int e = ctx->scan_op.reserve((size_t)P_pad * N * 32)) return e;
  int* d_max = reinterpret_cast<int*>(ctx->scal.as<float>() + SC_MMA_MAXCOUNT);
  TDR_CUDA(cudaMemsetAsync(d_max, 0, 8, ctx->stream));     // max count, "tensor-core kernel bailed out" flag
  {
    long long total = (long long)P_pad * N;
    k_build_scan_operand<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(
        ctx->scan_img.as<float>(), ctx->C, ctx->n_theta, ctx->n_r, P, P_pad, dev_shifts, n_shifts, S_pad, N,
        ctx->scan_op.as<uint4>(), d_max);
    count_launch(ctx);
    TDR_CUDA(cudaGetLastError());
    if (!grid_mode) {          // feeds the operand-format predictor of score_mma_i8.cu
      TDR_CUDA(cudaMemcpyAsync(ctx->scan_max_pin, d_max, 4, cudaMemcpyDeviceToHost, ctx->stream));
      TDR_CUDA(cudaEventRecord(ctx->scan_max_ev, ctx->stream));
      ctx->scan_max_pending = true;
    }
  }
  // counts above 2048 are not exact in fp16: checked ON THE DEVICE (sp.maxcount), the caller launches the guarded
  // CUDA-core kernel behind this one

  if (int e = build_perm(ctx, grid_mode, n_items)) return e;
  tdr::Particles& pt = ctx->part[ctx->cur];
  if (grid_mode) { if (int e = sync_const_tab_scaled(ctx, P, grid_scale, res)) return e; }
  else if (int e = sync_const_tab(ctx, P)) return e;
  ListParams sp; memset(&sp, 0, sizeof(sp));
  if (int e = build_map16(ctx, grid_mode ? ctx->grid_phase_log2 : 0, &sp.map16, &sp.geom)) return e;
  sp.resolution = ctx->resolution; sp.tab_scaled = grid_mode ? 1 : 0;
  sp.tab = ctx->tab.as<float2>(); sp.P = P; sp.P_pad = P_pad; sp.res = res;
  sp.bop = ctx->scan_op.as<uint4>();
  sp.perm = ctx->perm.as<int>();
  sp.n_shifts = n_shifts;
  sp.maxcount = d_max; sp.bailed = d_max + 1;
  if (grid_mode) {
    sp.n_work = n_items; sp.centers = ctx->grid_centers.as<float>(); sp.grid_scale = grid_scale;
    sp.costs = ctx->grid_costs.as<float>();
  } else {
    sp.n_work = ctx->n_uninit;
    sp.init_x = pt.init_x.as<float>(); sp.init_y = pt.init_y.as<float>(); sp.dx = pt.dx.as<float>(); sp.dy = pt.dy.as<float>();
    sp.theta = pt.theta.as<float>(); sp.scale = pt.scale.as<float>(); sp.have_init = pt.have_init.as<uint8_t>();
    sp.weights = ctx->weights.as<float>();
    sp.force_on_map = ctx->fp.force_on_map;
    sp.map_w = (float)ctx->cols * ctx->resolution; sp.map_h = (float)ctx->rows * ctx->resolution;
    sp.scale_gate = ctx->fp.fixed_scale < 0 ? 1 : 0;
    sp.scale_lo = pow(10.0, (double)ctx->fp.scale_log_min); sp.scale_hi = pow(10.0, (double)ctx->fp.scale_log_max);
    sp.regularization = ctx->fp.regularization;
    sp.thetas = ctx->d_search_thetas.as<float>();
  }
#define TDR_LAUNCH_LIST(IDX, NN, TT, RR, AA)                                                                            \
  do {                                                                                                                \
    using Cfg = ListCfg<NN, TT, RR, AA>;                                                                               \
    TDR_SMEM_OPTIN(ctx, OPTIN_LIST_BASE + IDX,                                                                        \
                   (k_score_mma_list<NN, TT, RR, AA>), Cfg::kSmem);                                                   \
    const long long nb = (sp.n_work + 128 * TT - 1) / (128 * TT);                                                     \
    long long cap = (long long)ctx->sm_count * (ctx->mma_ctas > 0 && ctx->mma_ctas < Cfg::kCtasPerSm ? ctx->mma_ctas : Cfg::kCtasPerSm); \
    if (ctx->mma_grid_cap > 0 && ctx->mma_grid_cap < cap) cap = ctx->mma_grid_cap;                                    \
    const int grid = (int)(nb < cap ? nb : cap);                                                                      \
    k_score_mma_list<NN, TT, RR, AA><<<grid, Cfg::kThreads, Cfg::kSmem, ctx->stream>>>(sp);                              \
  } while (0)
  const int cfg = ctx->mma_tiles * 10 + ctx->mma_split;
  if (ctx->mma_a_tmem) {
    switch (cfg) {
      case 21: TDR_LAUNCH_LIST(0, 96, 2, 1, true); break;
      case 12: TDR_LAUNCH_LIST(1, 96, 1, 2, true); break;
      case 14: TDR_LAUNCH_LIST(2, 96, 1, 4, true); break;
      default: TDR_LAUNCH_LIST(3, 96, 2, 2, true); break;     // 9.5 ms per 1e6 x 40 (T = 1, R = 4: 12.1; T = 2, R = 1: 15.6)
    }
  } else {
    switch (cfg) {
      case 41: TDR_LAUNCH_LIST(4, 96, 4, 1, false); break;
      case 21: TDR_LAUNCH_LIST(5, 96, 2, 1, false); break;
      case 22: TDR_LAUNCH_LIST(6, 96, 2, 2, false); break;
      case 11: TDR_LAUNCH_LIST(7, 96, 1, 1, false); break;
      case 14: TDR_LAUNCH_LIST(8, 96, 1, 4, false); break;
      default: TDR_LAUNCH_LIST(9, 96, 1, 2, false); break;
    }
  }
#undef TDR_LAUNCH_LIST
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  *used = true;
  return TDR_OK;
}

}  // namespace tdr
