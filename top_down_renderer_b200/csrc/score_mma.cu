// score_mma.cu — a7/a9/a10 as a tcgen05 gather-GEMM (theta search and exhaustive grid).
//
// Reference: TopDownMapPolar::getLocalMap (src/top_down_map_polar.cpp:21-53), StateParticle::getCostForRot
// (src/state_particle.cpp:112-155), StateParticle::computeWeight (:157-219).
//
// For a tile of 128 hypotheses the costs of ALL n_theta row shifts are matrix products accumulated over the
// lattice cells p = (theta, r):
//     Dcost[m, s] += sum_j A_p[m, j] * count_j[(theta + s) mod n_theta, r]          (j = class slot, hi and lo)
//     Dnorm[m, s] += known_p[m]      * tot    [(theta + s) mod n_theta, r]          (tot = class-summed count)
//               — the second product has ONE useful K slot per cell, so it is packed along K instead: the known
//               flags of 16 consecutive cells form one 128 x 16 operand (written by the gather threads, 2 bytes per
//               cell) and meet a precomputed 112 x 16 block of tot values in one MMA per 16 cells.
//   A_p[m, :] = the map record of hypothesis m at cell p: 16 fp16 = 8 hi + 8 lo halves of w_c * dist_c (class
//               weight folded into the map copy; hi + lo carries 22 mantissa bits), slot 7 hi = known.
//               One 32-byte record = one L2 sector = one K = 16 step of the MMA.
//   B         = a SLIDING WINDOW over the scan ring of radius bin r: the ring (n_theta + 112 rows of 16 fp16,
//               wrapped) sits in shared memory once per ring and the operand of cell theta is simply the
//               descriptor start address advanced by theta rows — no per-cell operand traffic at all.
// fp32 accumulation in TMEM (2 x 112 columns per tile).  cost[s] = 0.01 * Dcost[m, s] / Dnorm[m, s] (:136-154);
// the theta search picks its candidate shifts out of the n_theta columns in the epilogue.
//
// Warp roles (T tiles of 128 hypotheses per CTA, R gather threads per hypothesis row, persistent over batches):
//   gather warps   thread = (hypothesis row, 1 of R): computes the lattice pixel of its hypothesis for its stages
//                  (exact index math of tdr_math.cuh), pulls the 32-byte record with one 256-bit load and stores
//                  it into the K-major UMMA operand layout; thread 0 of each row runs the epilogue on its TMEM row.
//   loader warp    one lane streams the scan ring of the next radius bin and the tot blocks with cp.async.bulk.
//   MMA warp       one lane issues tcgen05.mma (cta_group::1, kind::f16, M = 128, N = 112, K = 16) once per
//                  (cell, tile) plus once per 16 cells for the normalisation; tcgen05.commit frees the stage / the
//                  ring slot / the tot slot / publishes the accumulators.
#include "mma_common.cuh"

namespace tdr {

// scan rings, per radius bin r: [kc = 2][x = n_theta + RING_N rows][8 halfs]; row x holds the class counts at angle
// x mod n_theta in the hi AND the lo slots (slot 7 zero).  K-major canonical layout with LBO = rows*16, SBO = 128,
// so the operand of cell theta starts theta*16 bytes into the ring.  One thread per (r, x).
// maxcount: device int, max class-summed count (fp16 integers are exact up to 2048).
static const int RING_N = 112;                     // accumulator columns per half (>= n_theta, multiple of 16)
static __global__ void k_build_rings(const float* __restrict__ img, int C, int n_theta, int n_r, uint4* __restrict__ out,
                              int* __restrict__ maxcount) {
  const int rows = n_theta + RING_N;
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n_r * rows) return;
  const int r = id / rows, x = id - r * rows;
  const int cell = r * n_theta + (x % n_theta);
  const int P = n_theta * n_r;
  unsigned short h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float tot = 0.f;
  for (int c = 0; c < C; c++) {
    float v = img[(size_t)c * P + cell];
    tot += v;
    h[c] = __half_as_ushort(__float2half_rn(v));
  }
  if (x < n_theta) atomicMax(maxcount, (int)tot);
  uint4 q;
  q.x = h[0] | ((uint32_t)h[1] << 16); q.y = h[2] | ((uint32_t)h[3] << 16);
  q.z = h[4] | ((uint32_t)h[5] << 16); q.w = h[6];                                   // slot 7 = 0
  uint4* ring = out + (size_t)r * rows * 2;       // 2 planes of `rows` uint4 per ring
  ring[x] = q;                    // kc 0 (hi slots)
  ring[rows + x] = q;             // kc 1 (lo slots see the same counts)
}

// normalisation operand, per group g of 16 consecutive cells p = 16 g + q: [kc = 2][n = RING_N][8 halfs] with
// element (n, q) = tot[(theta_p + n) mod n_theta, r_p]  (0 beyond the lattice or the n_theta shifts).
static __global__ void k_build_norm_op(const float* __restrict__ img, int C, int n_theta, int n_r, int n_groups,
                                uint4* __restrict__ out) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n_groups * RING_N) return;
  const int g = id / RING_N, n = id - g * RING_N;
  const int P = n_theta * n_r;
  unsigned short h[16];
#pragma unroll
  for (int q = 0; q < 16; q++) {
    const int p = g * 16 + q;
    float tot = 0.f;
    if (p < P && n < n_theta) {
      const int r = p / n_theta, th = p - r * n_theta;
      const int cell = r * n_theta + (th + n) % n_theta;
      for (int c = 0; c < C; c++) tot += img[(size_t)c * P + cell];
    }
    h[q] = __half_as_ushort(__float2half_rn(tot));
  }
  uint4 a, b;
  a.x = h[0] | ((uint32_t)h[1] << 16); a.y = h[2] | ((uint32_t)h[3] << 16); a.z = h[4] | ((uint32_t)h[5] << 16); a.w = h[6] | ((uint32_t)h[7] << 16);
  b.x = h[8] | ((uint32_t)h[9] << 16); b.y = h[10] | ((uint32_t)h[11] << 16); b.z = h[12] | ((uint32_t)h[13] << 16); b.w = h[14] | ((uint32_t)h[15] << 16);
  uint4* blk = out + (size_t)g * RING_N * 2;
  blk[n] = a;                     // kc 0: cells q = 0..7
  blk[RING_N + n] = b;            // kc 1: cells q = 8..15
}

// ------------------------------------------------------------------------------------------------
// the gather-GEMM
// ------------------------------------------------------------------------------------------------
struct MmaParams {
  const uint4* map16; GatherGeom geom; float resolution; int tab_scaled;
  int n_theta, n_r, P; float res;
  const uint4* rings; const uint4* norm_op; int n_groups;
  const int* perm; long long n_work;
  // particle mode
  const float *init_x, *init_y, *dx, *dy; float* theta; const float* scale; uint8_t* have_init; float* weights;
  int force_on_map; float map_w, map_h; int scale_gate; double scale_lo, scale_hi; float regularization;
  const float* thetas; const int32_t* shifts; int n_shifts;
  // grid mode
  const float* centers; float grid_scale; float* costs;
  // fused all-gather: every cost is stored into all ranks' full arrays (NVLink peer mappings) at this rank's rows
  float* cost_peers[TDR_MAX_PEERS]; int n_cost_peers; long long cost_row0;
  int peer_self;        // index of this rank's own array in cost_peers: destinations are visited from peer_self + 1 on
  int store_hint;       // 1: cost stores carry an L2 evict-first policy
  unsigned long long* grid_key;     // grid mode: (min cost, first global flat index) over everything this launch computes
  int identity_shifts;      // shifts[k] == k for all k and n_shifts % 4 == 0: vector stores of the cost rows
  const int* maxcount; int* bailed;     // device-side fp16 exactness precondition, see score_mma_list.cu
  int track_mode;           // particle mode: every particle takes the cost at ITS OWN heading's shift (rot_to_shift), no search
};

static const int MAX_RING_ROWS = 2 * RING_N;                  // n_theta <= RING_N
static const int RING_SLOT_BYTES = 2 * MAX_RING_ROWS * 16;    // 2 K chunks
static const int NB2 = 4, B2_BYTES = 2 * RING_N * 16;         // tot-block slots (one block per 16 cells)
// known-flag operand buffers (128 rows x 16 cells), one per 16-cell group in flight.  A gather thread may run kStages
// stages ahead of the stage the MMA warp has retired, and the last group of a batch is SHORT (P / G stages is not a
// multiple of the stages per group), so kStages + 1 consecutive stages can touch FOUR groups: the tail of one, the
// short one, a whole one of the next batch and the head of the one after.  With three buffers that fourth group
// overwrote the flags of the first before its normalisation MMA had read them (round 1: run-to-run differences of
// ~7e-5 in ~0.1 % of the grid costs whenever the long grid epilogue let the other thread sets run ahead —
// tools/determinism.py).  Needs kStages <= 2 x stages per group (asserted in MmaCfg).
static const int NA2 = 4, A2_TILE = 4096;
static const int CELLS_PER_GROUP = 16;
static const int OUT_STRIDE = RING_N;              // floats per staged cost row (16-byte multiple, >= n_theta)
// T = 128-hypothesis tiles per CTA; R = gather threads per hypothesis row (the R threads of a row take turns
// stage by stage, so the loads in flight per SM grow without growing the set of hypotheses — and their map
// footprint — that are in flight together)
// ATM (T = 1 only) = the gathered records and the known-flag operand live in TENSOR MEMORY (tcgen05.st; the MMA takes A
// from there): no operand stores through the L1 data pipe and no operand reads by the tensor core out of shared
// memory.  TMEM columns: [0, 224) accumulators | [224, 256) NA2 flag groups | [256, 512) 16 stages x 2 cells x 8.
// G = lattice cells per pipeline stage (2, or 4 on the tensor-memory path): the per-stage handshake of the single
// MMA-issuing warp (~90 instructions) is what bounds this kernel once the operands are off the L1 pipe, so a stage
// carries as many cells as the registers of the gather threads allow.
template <int T, int R, bool ATM, int G> struct MmaCfg {
  static const int kThreads = 128 * T * R + 64;
  static const int kTmemCols = ATM ? 512 : (T * 2 * RING_N <= 256 ? 256 : 512);
  static const int kRegs = G == 2 ? 96 : 144;
  static const int kByTmem = 512 / kTmemCols, kByRegs = 65536 / (kThreads * kRegs) < 1 ? 1 : 65536 / (kThreads * kRegs);
  static const int kCtasPerSm = kByTmem < kByRegs ? kByTmem : kByRegs;
  static const int kStageBytes = ATM ? 0 : G * T * A_TILE;
  static const int kA2Bytes = ATM ? 0 : NA2 * T * A2_TILE;
  // tensor-memory path: the epilogue stages every cost row in shared memory (OUT_STRIDE floats per row) and hands it
  // to the async proxy (cp.async.bulk shared -> global / peer) — the SM goes on with the next tile while rows drain
  static const int kOutBytes = ATM ? 128 * T * OUT_STRIDE * 4 : 0;
  static const int kFixed = 2 * RING_SLOT_BYTES + NB2 * B2_BYTES + kA2Bytes + 128 * T * R * 4 + 1024 + 4 * T * 32 * 17 * 4 + kOutBytes;
  static const int kBudget = (216 * 1024) / kCtasPerSm - 1280 - kFixed;
  // every gather thread keeps two of ITS stages in flight, i.e. spans 2R stages: leave twice that as slack
  static const int kStagesMax = 4 * R < 8 ? 8 : 4 * R;
  static const int kA2Col = T * 2 * RING_N, kACol = kA2Col + NA2 * T * 8, kACols = G * T * 8;   // ATM column map
  static const int kStagesSm = kBudget / (ATM ? 1 : kStageBytes) > kStagesMax ? kStagesMax : kBudget / (ATM ? 1 : kStageBytes);
  static const int kStages = ATM ? ((512 - kACol) / kACols > 16 ? 16 : (512 - kACol) / kACols) : kStagesSm;
  static const int kSmem = kStages * kStageBytes + kFixed;
  static_assert(!ATM || T == 1, "the tensor-memory operand path holds one tile per CTA");
  static_assert(kStages <= 2 * (CELLS_PER_GROUP / G), "NA2 = 4 flag buffers cover at most 2 x stages-per-group stages in flight");
  static_assert(G == 2 || (ATM && G == 4), "4-cell stages exist on the tensor-memory path only");
};

template <int T, int R, bool ATM, int G>
__global__ void __launch_bounds__(128 * T * R + 64, MmaCfg<T, R, ATM, G>::kCtasPerSm) k_score_mma(MmaParams sp) {
  using Cfg = MmaCfg<T, R, ATM, G>;
  if (*sp.maxcount > MMA_MAX_EXACT_COUNT) {          // grid-uniform, before anything is allocated
    if (blockIdx.x == 0 && threadIdx.x == 0) *sp.bailed = 1;
    return;
  }
  constexpr int STAGES_PER_GROUP = CELLS_PER_GROUP / G;
  constexpr int GW = 4 * T * R;        // gather warps
  constexpr int NS = Cfg::kStages;
  constexpr int ACC = 2 * RING_N;      // accumulator columns per tile: cost | norm
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sA = smem;                                         // [NS][G][T] tiles of A_TILE bytes
  unsigned char* sRing = smem + (size_t)NS * Cfg::kStageBytes;       // [2] ring slots
  unsigned char* sB2 = sRing + 2 * RING_SLOT_BYTES;                  // [NB2] tot blocks
  unsigned char* sA2 = sB2 + NB2 * B2_BYTES;                         // [NA2][T] known-flag tiles
  int* s_known = reinterpret_cast<int*>(sA2 + Cfg::kA2Bytes);        // [R][128 * T] known-cell counts
  // barriers: full[NS] empty[NS] ring_full[2] ring_empty[2] accum b2_full[NB2] b2_empty[NB2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_known + 128 * T * R);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * NS + 5 + 2 * NB2);
  short* s_inv = reinterpret_cast<short*>(s_tmem + 2);                  // [RING_N] shift -> first position in the list
  float* s_tile = reinterpret_cast<float*>(s_inv + RING_N + 8);         // [4T warps][32][17] epilogue transpose tiles
  float* s_out = s_tile + 4 * T * 32 * 17;                              // [128 T][OUT_STRIDE] staged cost rows (ATM)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + NS);
  const uint32_t bar_rfull = smem_u32(bars + 2 * NS), bar_rempty = smem_u32(bars + 2 * NS + 2), bar_accum = smem_u32(bars + 2 * NS + 4);
  const uint32_t bar_b2full = smem_u32(bars + 2 * NS + 5), bar_b2empty = smem_u32(bars + 2 * NS + 5 + NB2);

  if (warp == GW + 1) tmem_alloc(smem_u32(s_tmem), Cfg::kTmemCols);
  for (int q = tid; q < Cfg::kA2Bytes / 4; q += blockDim.x) reinterpret_cast<uint32_t*>(sA2)[q] = 0u;
  if (tid < RING_N) {
    int k = -1;
    for (int q = sp.n_shifts - 1; q >= 0; q--) if (sp.shifts[q] == tid) k = q;     // first occurrence wins
    s_inv[tid] = (short)k;
  }
  if (tid == 0) {
    for (int s = 0; s < NS; s++) { mbar_init(bar_full + 8 * s, 4 * T); mbar_init(bar_empty + 8 * s, 1); }
    for (int s = 0; s < 2; s++) { mbar_init(bar_rfull + 8 * s, 1); mbar_init(bar_rempty + 8 * s, 1); }
    for (int s = 0; s < NB2; s++) { mbar_init(bar_b2full + 8 * s, 1); mbar_init(bar_b2empty + 8 * s, 1); }
    mbar_init(bar_accum, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  if (ATM) {
    // flag columns start at zero: the cells past the lattice in the last group meet zero tot values, and 0 x garbage
    // must not be NaN
    if (warp < 4) {
      for (int q = 0; q < NA2 * T * 8; q++) tmem_st1(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(Cfg::kA2Col + q), 0u);
      tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  const long long per_batch = 128 * T;
  const long long n_batches = (sp.n_work + per_batch - 1) / per_batch;
  const int n_theta = sp.n_theta;
  const int K_ITERS = sp.P / G;                     // G divides n_theta: a stage never straddles two rings
  const int stages_per_ring = n_theta / G;
  const int ring_rows = n_theta + RING_N;
  const uint32_t ring_bytes = (uint32_t)ring_rows * 32;      // 2 planes x rows x 16 B
  uint32_t grp_it = 0;             // 16-cell group counter, continues across batches like `it`
  uint32_t it = 0;                 // pipeline iteration counter, continues across batches (same sequence in every role)
  uint32_t local_batch = 0;

  if (warp < GW) {
    // =========================== gather + epilogue ===========================
    const int sub = warp / (4 * T);                      // which of the R threads of a row this is
    const int t = (warp % (4 * T)) >> 2, m = tid & 127;
    const int row = tid % (128 * T);
    const unsigned char* map_bytes = reinterpret_cast<const unsigned char*>(sp.map16);
    for (long long batch = blockIdx.x; batch < n_batches; batch += gridDim.x, local_batch++) {
      const long long slot = batch * per_batch + row;
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
      int known_cnt = 0;
      const uint32_t grp_base = grp_it;

      // Each thread pulls the whole 32-byte record of ITS hypothesis with one 256-bit load (one sector, one L1
      // wavefront; measured 0.95 records/clk/SM from L2 against 0.42 for 2 x LDG.128 — tools/gather_bench.cu).
      auto load_stage = [&](int k, uint4 (&rec)[G][2]) {
#pragma unroll
        for (int g = 0; g < G; g++) {
          const int p = k * G + g;
          const float2 tb = c_tab[p];
          float vy, vx;
          if (sp.tab_scaled) { vy = TDR_FADD(tb.x, oy); vx = TDR_FADD(tb.y, ox); }            // grid: the table holds (tab * scale) * res
          else { vy = TDR_FADD(TDR_FMUL(TDR_FMUL(tb.x, sc), sp.res), oy); vx = TDR_FADD(TDR_FMUL(TDR_FMUL(tb.y, sc), sp.res), ox); }
          uint32_t off = lattice_record(vy, vx, sp.geom);
          if (!active) off = sp.geom.zero_rec;
          ldg256(map_bytes + (size_t)off * 32, rec[g][0], rec[g][1]);
        }
      };
      auto store_stage = [&](int k, const uint4 (&rec)[G][2]) {
        const uint32_t iter = it + (uint32_t)k;
        const uint32_t st = iter % NS, ph = (iter / NS) & 1u;
        mbar_wait(bar_empty + 8 * st, ph ^ 1u);
        // the known flags as one K = 16 operand per 16 cells: this stage owns two adjacent halfs of row m
        uint32_t word[G / 2];
#pragma unroll
        for (int g = 0; g < G; g += 2)
          word[g / 2] = (rec[g][0].w & 0xffff0000u ? 0x00003C00u : 0u) | (rec[g + 1][0].w & 0xffff0000u ? 0x3C000000u : 0u);
        const uint32_t gq = (uint32_t)k / STAGES_PER_GROUP, q0 = ((uint32_t)k % STAGES_PER_GROUP) * G;
        const uint32_t gslot = (grp_base + gq) % NA2;
#pragma unroll
        for (int g = 0; g < G; g++) known_cnt += (rec[g][0].w >> 16) != 0u ? 1 : 0;    // slot 7 hi = known (1.0)
        if (ATM) {
          // row m = TMEM lane m (this warp's quarter); a record = 8 columns (16 fp16 in K order), a flag pair = 1 column
          const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll
          for (int g = 0; g < G; g++)
            tmem_st8(lane_base + (uint32_t)(Cfg::kACol + st * Cfg::kACols + (g * T + t) * 8), rec[g][0], rec[g][1]);
          if (G == 2) tmem_st1(lane_base + (uint32_t)(Cfg::kA2Col + (gslot * T + t) * 8 + (q0 >> 1)), word[0]);
          else tmem_st2(lane_base + (uint32_t)(Cfg::kA2Col + (gslot * T + t) * 8 + (q0 >> 1)), word[0], word[G / 2 - 1]);
          tmem_wait_st();
          tc_fence_before();
        } else {
          unsigned char* base = sA + (size_t)st * Cfg::kStageBytes + (size_t)t * A_TILE + (size_t)m * 16;
#pragma unroll
          for (int g = 0; g < G; g++) {
            *reinterpret_cast<uint4*>(base + (size_t)g * T * A_TILE) = rec[g][0];            // K chunk 0: hi halves
            *reinterpret_cast<uint4*>(base + (size_t)g * T * A_TILE + A_LBO) = rec[g][1];    // K chunk 1: lo halves
          }
          *reinterpret_cast<uint32_t*>(sA2 + (size_t)(gslot * T + t) * A2_TILE + (q0 >> 3) * 2048 + (size_t)m * 16 + (q0 & 7) * 2) = word[0];
          fence_proxy_async();
        }
        if (k + R >= K_ITERS) s_known[sub * 128 * T + row] = known_cnt;     // this thread's last stage of the batch
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full + 8 * st);
      };

      uint4 ra[G][2], rb[G][2];
      if (sub < K_ITERS) load_stage(sub, ra);
#pragma unroll 1
      for (int k = sub; k < K_ITERS; k += 2 * R) {       // this thread's stages: sub, sub + R, ... (two in flight)
        if (k + R < K_ITERS) load_stage(k + R, rb);
        store_stage(k, ra);
        if (k + 2 * R < K_ITERS) load_stage(k + 2 * R, ra);
        if (k + R < K_ITERS) store_stage(k + R, rb);
      }
      it += K_ITERS;
      grp_it += (uint32_t)sp.n_groups;
      if (sub != 0) continue;                            // thread 0 of each row owns the epilogue

      // ---- epilogue: this hypothesis' accumulator row
      mbar_wait(bar_accum, local_batch & 1u);
      tc_fence_after();
      int known = 0;
#pragma unroll
      for (int q = 0; q < R; q++) known += s_known[q * 128 * T + row];
      const bool unknown = (double)TDR_FDIV((float)known, (float)sp.P) < 0.5;                // :117-120
      const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(t * ACC);
      float best = 3.402823466e+38f;                                                         // :193-204
      int best_k = 0x7fffffff;
      // tracking: the one column of this particle's heading (state_particle.cpp:207-210); NaN stays NaN there
      const int my_shift = (sp.track_mode && i >= 0) ? rot_to_shift(sp.theta[i], n_theta) : -1;
      float my_cost = 0.f;
      uint32_t vc[16], vn[16];
      const bool async_rows = ATM && sp.identity_shifts && (sp.costs || sp.n_cost_peers);   // CTA-uniform
      if (async_rows) { bulk_wait_read(); __syncwarp(); }   // the previous tile's stores (issued by the run heads of this warp) have read the staging rows
#pragma unroll 1
      for (int ch = 0; ch * 16 < n_theta; ch++) {
        tmem_ld16(trow + (uint32_t)(ch * 16), vc);
        tmem_ld16(trow + (uint32_t)(RING_N + ch * 16), vn);
        tmem_wait_ld();
        float cst[16];
#pragma unroll
        for (int j = 0; j < 16; j++) {
          const int s = ch * 16 + j;
          const int k = s < n_theta ? s_inv[s] : -1;       // position of shift s in the candidate list (-1: not asked)
          cst[j] = unknown ? __int_as_float(0x7fc00000)
                           : TDR_FDIV(TDR_FMUL(__uint_as_float(vc[j]), 0.01f), __uint_as_float(vn[j]));   // :137,154
          // first strict minimum in LIST order == lexicographic minimum of (cost, list position)
          if (k >= 0 && (cst[j] < best || (cst[j] == best && k < best_k))) { best = cst[j]; best_k = k; }
          if (s == my_shift) my_cost = cst[j];
        }
        if ((sp.costs || sp.n_cost_peers)) {          // warp-uniform
          const int n_dst = sp.n_cost_peers ? sp.n_cost_peers : 1;
          const long long grow = (sp.n_cost_peers ? sp.cost_row0 + i : i) * sp.n_shifts;
          if (ATM && sp.identity_shifts) {
            // all shifts in order: column s is list position s.  Stage this row; the bulk stores go out after the loop.
            // rows are staged DENSELY (n_shifts floats apart) so that consecutive hypotheses form one contiguous block
            float* mine = s_out + (size_t)row * sp.n_shifts + ch * 16;
#pragma unroll
            for (int q = 0; q < 4; q++)
              if (ch * 16 + 4 * q + 3 < sp.n_shifts)
                *reinterpret_cast<float4*>(mine + 4 * q) = make_float4(cst[4 * q], cst[4 * q + 1], cst[4 * q + 2], cst[4 * q + 3]);
          } else
          if (sp.identity_shifts) {
            // all shifts in order: column s is list position s.  The warp's 32 rows x 16 columns go through a
            // shared-memory tile so that four lanes write one row's 64 contiguous bytes (8 rows per store
            // instruction) — the difference between usable and useless NVLink packets for the peer copies.
            float* tile = s_tile + (size_t)(warp % (4 * T)) * (32 * 17);
#pragma unroll
            for (int j = 0; j < 16; j++) tile[lane * 17 + j] = cst[j];
            __syncwarp();
            const int piece = lane & 3;
            const bool col_ok = ch * 16 + piece * 4 + 3 < sp.n_shifts;
#pragma unroll
            for (int q = 0; q < 4; q++) {
              const int r = q * 8 + (lane >> 2);
              const long long ir = __shfl_sync(0xffffffffu, i, r);
              const float* src = tile + r * 17 + piece * 4;
              const float4 v = make_float4(src[0], src[1], src[2], src[3]);
              if (ir >= 0 && col_ok) {
                const long long at = (sp.n_cost_peers ? sp.cost_row0 + ir : ir) * sp.n_shifts + ch * 16 + piece * 4;
#pragma unroll 1
                for (int d = 0; d < n_dst; d++)
                  *reinterpret_cast<float4*>((sp.n_cost_peers ? sp.cost_peers[d] : sp.costs) + at) = v;
              }
            }
            __syncwarp();
          } else {
#pragma unroll
            for (int j = 0; j < 16; j++) {
              const int s = ch * 16 + j;
              const int k = s < n_theta ? s_inv[s] : -1;
              if (k >= 0 && i >= 0) {
#pragma unroll 1
                for (int d = 0; d < n_dst; d++) (sp.n_cost_peers ? sp.cost_peers[d] : sp.costs)[grow + k] = cst[j];
              }
            }
          }
        }
      }
      if (async_rows) {
        // one bulk store of the whole row per destination (own array and every peer's, NVLink): issued here, executed
        // by the async proxy while this CTA gathers the next tile
        // The hypotheses of a warp are mostly CONSECUTIVE (the spatial order keeps lattice rows together), so their cost
        // rows are contiguous at the destination too: the first lane of every run stores the whole run with one bulk
        // copy per destination (up to 32 rows = 12.8 KB) instead of one 400-byte copy per row — 30 x fewer operations
        // in the SM's bulk-copy queue, which the scan-ring loads of the next tile share, and full-size NVLink packets.
        fence_proxy_async();
        __syncwarp();
        {
          const long long prev_i = __shfl_up_sync(0xffffffffu, i, 1);
          const bool cont = lane > 0 && i >= 0 && prev_i >= 0 && i == prev_i + 1;
          const uint32_t breaks = __ballot_sync(0xffffffffu, !cont);
          if (i >= 0 && !cont) {
            const uint32_t later = lane == 31 ? 0u : (breaks & ~((2u << lane) - 1u));
            const int len = (later ? __ffs(later) - 1 : 32) - lane;
            const long long at = (sp.n_cost_peers ? sp.cost_row0 + i : i) * sp.n_shifts;
            const uint32_t src = smem_u32(s_out + (size_t)row * sp.n_shifts);
            const int n_dst = sp.n_cost_peers ? sp.n_cost_peers : 1;
            // destinations in an order that differs per rank and per CTA: with every rank walking 0, 1, 2 ... all eight
            // GPUs would write to the same peer at the same moment (one ingress port 7x oversubscribed, six idle)
            int d = (sp.n_cost_peers && sp.peer_self >= 0) ? (sp.peer_self + 1 + (int)(blockIdx.x % (unsigned)n_dst)) % n_dst : 0;
            const uint32_t bytes = (uint32_t)(len * sp.n_shifts) * 4u;
#pragma unroll 1
            for (int j = 0; j < n_dst; j++) {
              float* dst = (sp.n_cost_peers ? sp.cost_peers[d] : sp.costs) + at;
              if (sp.store_hint) bulk_s2g_hint(dst, src, bytes, l2_policy_evict_first());
              else bulk_s2g(dst, src, bytes);
              if (++d == n_dst) d = 0;
            }
          }
        }
        bulk_commit();
      }
      if (sp.grid_key) {
        // fold this row's first minimum into the launch-wide (cost, flat index) key — k_grid_best's format
        unsigned long long key = ~0ull;
        if (i >= 0 && best_k != 0x7fffffff) {
          uint32_t u = __float_as_uint(best);
          u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
          key = ((unsigned long long)u << 32) | (unsigned long long)(uint32_t)((sp.cost_row0 + i) * sp.n_shifts + best_k);
        }
        for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, key, o); if (t < key) key = t; }
        if (lane == 0 && key != ~0ull) atomicMin(sp.grid_key, key);
      }
      if (i >= 0 && !sp.centers && sp.track_mode) {
        sp.weights[i] = gated ? 0.f : (float)(1.0 / (double)TDR_FADD(my_cost, sp.regularization));   // :212
      } else if (i >= 0 && !sp.centers) {
        if (gated) sp.weights[i] = 0.f;
        else {
          sp.theta[i] = best_k == 0x7fffffff ? 0.f : sp.thetas[best_k];
          sp.have_init[i] = 1;
          sp.weights[i] = (float)(1.0 / (double)TDR_FADD(best, sp.regularization));             // :212
        }
      }
      tc_fence_before();           // TMEM reads are done before the next batch's first full-barrier arrive
    }
    if (ATM) bulk_wait_all();      // this thread's row stores have landed (own memory and peers) before the kernel ends
  } else if (warp == GW) {
    // =========================== scan-ring / tot-block loader ===========================
    // whole warp, warp-uniform control flow; the elected lane issues the copies.  Order = the order in which the MMA
    // issuer consumes them: the ring of a radius bin before the tot block of the group in which its first cell lies.
    const bool leader = elect_one();
    uint32_t rsl = 0, rph = 0, bsl = 0, bph = 0;          // ring slot / parity, tot-block slot / parity
    for (long long batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
      int next_ring = 0, ring_first = 0;                  // next ring to load and its first stage
      for (int g = 0; g < sp.n_groups; g++) {
        while (next_ring < sp.n_r && ring_first <= g * STAGES_PER_GROUP + STAGES_PER_GROUP - 1) {
          mbar_wait(bar_rempty + 8 * rsl, rph ^ 1u);
          if (leader) {
            mbar_expect_tx(bar_rfull + 8 * rsl, ring_bytes);
            bulk_g2s(smem_u32(sRing) + rsl * RING_SLOT_BYTES,
                     reinterpret_cast<const unsigned char*>(sp.rings) + (size_t)next_ring * ring_bytes, ring_bytes, bar_rfull + 8 * rsl);
          }
          __syncwarp();
          rsl ^= 1u; if (rsl == 0) rph ^= 1u;
          next_ring++; ring_first += stages_per_ring;
        }
        mbar_wait(bar_b2empty + 8 * bsl, bph ^ 1u);
        if (leader) {
          mbar_expect_tx(bar_b2full + 8 * bsl, B2_BYTES);
          bulk_g2s(smem_u32(sB2) + bsl * B2_BYTES, reinterpret_cast<const unsigned char*>(sp.norm_op) + (size_t)g * B2_BYTES,
                   B2_BYTES, bar_b2full + 8 * bsl);
        }
        __syncwarp();
        if (++bsl == NB2) { bsl = 0; bph ^= 1u; }
      }
    }
  } else {
    // =========================== MMA issuer ===========================
    // whole warp, warp-uniform control flow, incremental slot / parity counters; the elected lane issues.
    // instruction descriptor: D = f32, A = B = f16, both K-major, N = 112, M = 128
    const bool leader = elect_one();
    const uint32_t idesc = (1u << 4) | ((uint32_t)(RING_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t plane = (uint32_t)ring_rows * 16;           // LBO of the ring operand
    const uint32_t sA_u = smem_u32(sA), sA2_u = smem_u32(sA2);
    // pin the addresses this loop lives on in registers: left to itself ptxas re-derives the shared-window base
    // (S2R CgaCtaId + LEA) in every iteration of the critical path
    uint32_t sRing_u = smem_u32(sRing), sB2_u = smem_u32(sB2), q_full = bar_full, q_empty = bar_empty, q_rfull = bar_rfull,
             q_rempty = bar_rempty, q_accum = bar_accum, q_b2full = bar_b2full, q_b2empty = bar_b2empty, q_tmem = tmem_base;
    asm volatile("" : "+r"(sRing_u), "+r"(sB2_u), "+r"(q_full), "+r"(q_empty), "+r"(q_rfull), "+r"(q_rempty), "+r"(q_accum),
                 "+r"(q_b2full), "+r"(q_b2empty), "+r"(q_tmem));
    if (ATM && T == 1) {
      // the tensor-memory path: every per-stage quantity is a RUNNING value (barrier addresses, the stage's TMEM
      // columns, the window descriptor), advanced by an add and wrapped by a compare — no multiplies, no re-derivation
      uint32_t ph = 0, rsl = 0, rph = 0, bph = 0;
      uint32_t full_a = q_full, empty_a = q_empty;                       // + 8 per stage, wrap after NS
      uint32_t a_cols = q_tmem + (uint32_t)Cfg::kACol;                   // + kACols per stage, wrap after NS
      uint32_t b2full_a = q_b2full, b2empty_a = q_b2empty, b2_addr = sB2_u;   // + 8 / + B2_BYTES per group, wrap after NB2
      uint32_t a2_cols = q_tmem + (uint32_t)Cfg::kA2Col;                 // + 8 per group, wrap after NA2
      const uint32_t full_end = q_full + 8 * NS, a_end = q_tmem + (uint32_t)(Cfg::kACol + NS * Cfg::kACols);
      const uint32_t b2full_end = q_b2full + 8 * NB2, a2_end = q_tmem + (uint32_t)(Cfg::kA2Col + NA2 * 8);
      for (long long batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
        uint64_t bdesc = 0;                                              // window of the stage's first cell (+ G per stage)
        uint32_t ring_empty_a = 0;
        int sr = 0, sg = 0;
        uint32_t acc = 0, acc_norm = 0;
        for (int k = 0; k < K_ITERS; k++) {
          if (sr == 0) {
            mbar_wait(q_rfull + 8 * rsl, rph);
            bdesc = umma_desc(sRing_u + rsl * RING_SLOT_BYTES, plane, 128);
            ring_empty_a = q_rempty + 8 * rsl;
            rsl ^= 1u; if (rsl == 0) rph ^= 1u;
          }
          mbar_wait(full_a, ph);
          const bool group_done = sg == STAGES_PER_GROUP - 1 || k == K_ITERS - 1;
          if (group_done) mbar_wait(b2full_a, bph);
          tc_fence_after();
          if (leader) {
            umma_stage_ts<G>(q_tmem, a_cols, bdesc, idesc, acc, 1u);
            if (group_done) {
              // normalisation: the known flags of this group's cells against the tot block
              umma_f16_ts(q_tmem + (uint32_t)RING_N, a2_cols, umma_desc(b2_addr, RING_N * 16, 128), idesc, acc_norm);
            }
            umma_commit(empty_a);                   // implies tcgen05.fence::before_thread_sync
            if (group_done) umma_commit(b2empty_a);
            if (sr == stages_per_ring - 1) umma_commit(ring_empty_a);   // ring slot free once these MMAs retire
            if (k == K_ITERS - 1) umma_commit(q_accum);
          }
          __syncwarp();
          acc = 1u;
          bdesc += (uint64_t)G;
          if (group_done) {
            sg = 0; acc_norm = 1u;
            b2full_a += 8; b2empty_a += 8; b2_addr += B2_BYTES;
            if (b2full_a == b2full_end) { b2full_a = q_b2full; b2empty_a = q_b2empty; b2_addr = sB2_u; bph ^= 1u; }
            a2_cols += 8; if (a2_cols == a2_end) a2_cols = q_tmem + (uint32_t)Cfg::kA2Col;
          } else sg++;
          if (++sr == stages_per_ring) sr = 0;
          full_a += 8; empty_a += 8; a_cols += (uint32_t)Cfg::kACols;
          if (full_a == full_end) { full_a = q_full; empty_a = q_empty; a_cols = q_tmem + (uint32_t)Cfg::kACol; ph ^= 1u; }
        }
      }
    } else {
    uint32_t st = 0, ph = 0;                                   // stage slot / parity (continue across batches)
      uint32_t rsl = 0, rph = 0, bsl = 0, bph = 0, asl = 0;      // ring, tot-block and flag-group slots
      for (long long batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
        uint64_t ring_desc = 0;
        uint32_t cur_rsl = 0;
        int sr = 0, sg = 0, grp = 0;                             // stage within the ring / within the 16-cell group; group
        for (int k = 0; k < K_ITERS; k++) {
          if (sr == 0) {
            mbar_wait(q_rfull + 8 * rsl, rph);
            ring_desc = umma_desc(sRing_u + rsl * RING_SLOT_BYTES, plane, 128);
            cur_rsl = rsl;
            rsl ^= 1u; if (rsl == 0) rph ^= 1u;
          }
          mbar_wait(q_full + 8 * st, ph);
          const bool group_done = sg == STAGES_PER_GROUP - 1 || k == K_ITERS - 1;
          if (group_done) mbar_wait(q_b2full + 8 * bsl, bph);
          tc_fence_after();
          if (leader) {
  #pragma unroll
            for (int g = 0; g < G; g++) {
              const uint64_t bcnt = ring_desc + (uint64_t)(uint32_t)(sr * G + g);      // window start advances 16 B per cell
  #pragma unroll
              for (int tt = 0; tt < T; tt++) {
                if (ATM)
                  umma_f16_ts(q_tmem + (uint32_t)(tt * ACC), q_tmem + (uint32_t)(Cfg::kACol + (g * T + tt) * 8) + st * Cfg::kACols,
                              bcnt, idesc, (k > 0 || g > 0) ? 1u : 0u);
                else
                  umma_f16(q_tmem + (uint32_t)(tt * ACC), umma_desc(sA_u + st * Cfg::kStageBytes + (g * T + tt) * A_TILE, A_LBO, 128),
                           bcnt, idesc, (k > 0 || g > 0) ? 1u : 0u);
              }
            }
            if (group_done) {
              // normalisation: the known flags of this group's cells against the tot block
              const uint64_t btot = umma_desc(sB2_u + bsl * B2_BYTES, RING_N * 16, 128);
  #pragma unroll
              for (int tt = 0; tt < T; tt++) {
                if (ATM)
                  umma_f16_ts(q_tmem + (uint32_t)(tt * ACC + RING_N), q_tmem + (uint32_t)(Cfg::kA2Col + tt * 8) + asl * (T * 8),
                              btot, idesc, grp > 0 ? 1u : 0u);
                else
                  umma_f16(q_tmem + (uint32_t)(tt * ACC + RING_N), umma_desc(sA2_u + (asl * T + tt) * A2_TILE, 2048, 128), btot, idesc,
                           grp > 0 ? 1u : 0u);
              }
            }
            umma_commit(q_empty + 8 * st);          // implies tcgen05.fence::before_thread_sync
            if (group_done) umma_commit(q_b2empty + 8 * bsl);
            if (sr == stages_per_ring - 1) umma_commit(q_rempty + 8 * cur_rsl);   // ring slot free once these MMAs retire
            if (k == K_ITERS - 1) umma_commit(q_accum);
          }
          __syncwarp();
          if (group_done) {
            sg = 0; grp++;
            if (++bsl == NB2) { bsl = 0; bph ^= 1u; }
            if (++asl == NA2) asl = 0;
          } else sg++;
          if (++sr == stages_per_ring) sr = 0;
          if (++st == NS) { st = 0; ph ^= 1u; }
        }
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
static bool mma_usable(tdr_ctx* ctx, const int32_t* host_shifts, int n_shifts) {
  if (ctx->score_impl == 1) return false;
  const int n_theta = ctx->n_theta;
  if (n_theta > RING_N || (n_theta % 2) != 0 || ctx->n_theta * ctx->n_r > MMA_TAB_MAX) return false;
  if (n_shifts < 1 || n_shifts > TDR_MAX_SHIFTS) return false;
  bool seen[RING_N] = {};
  for (int k = 0; k < n_shifts; k++) {              // every candidate must be a distinct row shift in [0, n_theta)
    const int s = host_shifts[k];
    if (s < 0 || s >= n_theta || seen[s]) return false;
    seen[s] = true;
  }
  for (int c = 0; c < ctx->C; c++) {
    float w = ctx->fp.class_weights[c];
    if (!(w >= 0.f) || w * 50.f > 60000.f) return false;     // fp16 range of w_c * dist_c (dist <= 50)
  }
  return true;
}

// returns TDR_OK and sets *used = true when the tensor-core path ran; *used = false -> caller falls back
int score_mma(tdr_ctx* ctx, float res, bool grid_mode, long long n_items, float grid_scale, const int32_t* dev_shifts,
              const int32_t* host_shifts, int n_shifts, bool* used, bool track) {
  *used = false;
  if (!mma_usable(ctx, host_shifts, n_shifts)) return TDR_OK;
  const int P = ctx->n_theta * ctx->n_r;
  const int ring_rows = ctx->n_theta + RING_N;
  // ---- scan rings (+ max count check: fp16 integers are exact up to 2048)
  const int n_groups = (P + CELLS_PER_GROUP - 1) / CELLS_PER_GROUP;
  const size_t ring_total = (size_t)ctx->n_r * ring_rows * 32;
  if (int e = ctx->scan_op.reserve(ring_total + (size_t)n_groups * B2_BYTES)) return e;
  uint4* d_norm_op = reinterpret_cast<uint4*>(ctx->scan_op.as<unsigned char>() + ring_total);
  int* d_max = reinterpret_cast<int*>(ctx->scal.as<float>() + SC_MMA_MAXCOUNT);
  TDR_CUDA(cudaMemsetAsync(d_max, 0, 8, ctx->stream));     // max count, "tensor-core kernel bailed out" flag
  {
    const int total = ctx->n_r * ring_rows;
    k_build_rings<<<(total + 255) / 256, 256, 0, ctx->stream>>>(ctx->scan_img.as<float>(), ctx->C, ctx->n_theta, ctx->n_r,
                                                                ctx->scan_op.as<uint4>(), d_max);
    k_build_norm_op<<<(n_groups * RING_N + 255) / 256, 256, 0, ctx->stream>>>(ctx->scan_img.as<float>(), ctx->C, ctx->n_theta,
                                                                               ctx->n_r, n_groups, d_norm_op);
    count_launch(ctx, 2);
    TDR_CUDA(cudaGetLastError());
  }
  // counts above 2048 are not exact in fp16: checked ON THE DEVICE (sp.maxcount), the caller launches the guarded
  // CUDA-core kernel behind this one

  if (int e = build_perm(ctx, grid_mode, n_items, false, track)) return e;
  tdr::Particles& pt = ctx->part[ctx->cur];
  if (grid_mode) { if (int e = sync_const_tab_scaled(ctx, P, grid_scale, res)) return e; }
  else if (int e = sync_const_tab(ctx, P)) return e;
  MmaParams sp; memset(&sp, 0, sizeof(sp));
  if (int e = build_map16(ctx, grid_mode ? ctx->grid_phase_log2 : 0, &sp.map16, &sp.geom)) return e;
  sp.resolution = ctx->resolution; sp.tab_scaled = grid_mode ? 1 : 0;
  sp.n_theta = ctx->n_theta; sp.n_r = ctx->n_r; sp.P = P; sp.res = res;
  sp.rings = ctx->scan_op.as<uint4>(); sp.norm_op = d_norm_op; sp.n_groups = n_groups;
  sp.perm = ctx->perm.as<int>();
  sp.shifts = dev_shifts; sp.n_shifts = n_shifts;
  sp.maxcount = d_max; sp.bailed = d_max + 1;
  if (grid_mode) {
    sp.n_work = n_items; sp.centers = ctx->grid_centers.as<float>(); sp.grid_scale = grid_scale;
    sp.costs = grid_costs_ptr(ctx);
    sp.n_cost_peers = ctx->grid_n_peers; sp.cost_row0 = ctx->grid_n_peers ? ctx->grid_peer_row0 : 0;
    if ((long long)(sp.cost_row0 + n_items) * n_shifts < (1ll << 32)) {
      if (int e = ctx->grid_key.reserve(8)) return e;
      TDR_CUDA(cudaMemsetAsync(ctx->grid_key.p, 0xff, 8, ctx->stream));
      sp.grid_key = ctx->grid_key.as<unsigned long long>();
    }
    sp.identity_shifts = (n_shifts % 4 == 0) ? 1 : 0;
    for (int k = 0; k < n_shifts; k++) if (host_shifts[k] != k) sp.identity_shifts = 0;
    for (int d = 0; d < ctx->grid_n_peers; d++) sp.cost_peers[d] = ctx->grid_peers[d];
    sp.peer_self = 0;
    for (int d = 0; d < ctx->grid_n_peers; d++) if (ctx->grid_peers[d] == ctx->grid_full.as<float>()) sp.peer_self = d;
    if (const char* e = getenv("TDR_GRID_ROTATE")) { if (!atoi(e)) sp.peer_self = -1; }       // diagnostic: every rank walks 0, 1, 2 ...
    sp.store_hint = ctx->grid_store_hint;
    if (ctx->grid_n_peers && ctx->grid_self_only) { sp.n_cost_peers = 1; sp.cost_peers[0] = ctx->grid_full.as<float>(); }   // diagnostic: no NVLink stores
  } else {
    sp.n_work = track ? pt.n - ctx->n_uninit : ctx->n_uninit;
    sp.track_mode = track ? 1 : 0;
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
#define TDR_LAUNCH_MMA(IDX, TT, RR, AA, GG)                                                                            \
  do {                                                                                                                \
    using Cfg = MmaCfg<TT, RR, AA, GG>;                                                                               \
    TDR_SMEM_OPTIN(ctx, OPTIN_RING_BASE + IDX,                                                                        \
                   (k_score_mma<TT, RR, AA, GG>), Cfg::kSmem);                                                        \
    const long long nb = (sp.n_work + 128 * TT - 1) / (128 * TT);                                                     \
    long long cap = (long long)ctx->sm_count * (ctx->mma_ctas > 0 && ctx->mma_ctas < Cfg::kCtasPerSm ? ctx->mma_ctas : Cfg::kCtasPerSm); \
    if (ctx->mma_grid_cap > 0 && ctx->mma_grid_cap < cap) cap = ctx->mma_grid_cap;                                    \
    const int grid = (int)(nb < cap ? nb : cap);                                                                      \
    k_score_mma<TT, RR, AA, GG><<<grid, Cfg::kThreads, Cfg::kSmem, ctx->stream>>>(sp);                                \
  } while (0)
  // tiles * 10 + threads per row (+ 100: operands in tensor memory, + 400: and 4-cell stages, needs n_theta % 4 == 0)
  int rcfg = ctx->mma_ring_cfg;
  if (rcfg >= 400 && ctx->n_theta % 4 != 0) rcfg = 114;
  switch (rcfg) {
    case 21: TDR_LAUNCH_MMA(0, 2, 1, false, 2); break;
    case 22: TDR_LAUNCH_MMA(1, 2, 2, false, 2); break;
    case 11: TDR_LAUNCH_MMA(2, 1, 1, false, 2); break;
    case 14: TDR_LAUNCH_MMA(3, 1, 4, false, 2); break;
    case 12: TDR_LAUNCH_MMA(4, 1, 2, false, 2); break;
    case 112: TDR_LAUNCH_MMA(5, 1, 2, true, 2); break;
    case 114: TDR_LAUNCH_MMA(6, 1, 4, true, 2); break;
    case 412: TDR_LAUNCH_MMA(7, 1, 2, true, 4); break;
    default: TDR_LAUNCH_MMA(8, 1, 3, true, 4); break;
  }
#undef TDR_LAUNCH_MMA
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  ctx->grid_key_valid = grid_mode && sp.grid_key != nullptr;
  *used = true;
  return TDR_OK;
}

}  // namespace tdr
