// tdr_ctx.cuh — internal context of libtdr_b200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/tdr.h"

namespace tdr {

void set_error(const char* fmt, ...);

#define TDR_CUDA(call)                                                                       \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      tdr::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return TDR_ECUDA;                                                                      \
    }                                                                                        \
  } while (0)

#define TDR_REQUIRE(cond, code, ...)        \
  do {                                      \
    if (!(cond)) {                          \
      tdr::set_error(__VA_ARGS__);          \
      return (code);                        \
    }                                       \
  } while (0)

// growable device buffer
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return TDR_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return TDR_ECUDA; }
    cap = bytes;
    return TDR_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// pinned host staging buffer (H2D / D2H without a pageable bounce)
struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return TDR_OK;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMallocHost(&p, bytes);
    if (e != cudaSuccess) { set_error("cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e)); return TDR_ECUDA; }
    cap = bytes;
    return TDR_OK;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// Device map pixel: 8 fp32 slots (32 B = one sector): slots [0, C) hold the class distance
// fields exactly as the reference's class_maps_ hold them after computeDists, slot 7 holds
// known = 1 - class_mask_ (1.0f / 0.0f).  Row-major over (row = y, col = x), x fastest.
struct MapPixel { float v[8]; };

struct Particles {   // SoA mirror of State (+ last_dist_, weight_)
  DevBuf init_x, init_y, dx, dy, theta, scale, have_init /*u8*/, last_dist;
  int64_t n = 0;
  int reserve(int64_t cap);
  void release();
};

}  // namespace tdr

struct tdr_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  int64_t launches = 0;

  // ---- map
  int rows = 0, cols = 0, C = 0;
  float resolution = 1.f;
  bool have_map = false;
  tdr::DevBuf map_px;        // rows*cols MapPixel
  tdr::DevBuf seedbits;      // rows*cols uint8: bit c set where class c is present (binary layer == 0)
  bool have_seeds = false;
  tdr::DevBuf edt_g;         // C planes rows*cols uint8 vertical distances (scratch)
  tdr::DevBuf scratch;       // generic scratch (layer staging, partial histograms, ...)
  tdr::DevBuf scratch2;

  // tensor-core score path (score_mma.cu): class-weighted fp16 hi/lo copy of the map, scan operand, binning
  tdr::DevBuf map16;         // rows*cols x 32 B
  bool map16_valid = false;
  // second copy for lattices of centres (exhaustive grid): every map row split into 2^k PHASE rows (x mod 2^k), so
  // that centres 2^k px apart read CONSECUTIVE records — 8 full 128-byte lines per warp load instead of one sector
  // out of each of 32 lines (measured 1.67 against 0.82 records/clk/SM, tools/gather_bench.cu patterns 7 / 4)
  tdr::DevBuf geo_planar;    // the 2 geometric distance layers (getGeoRasterMap + computeDists), col-major, built on demand
  bool geo_valid = false;
  tdr::DevBuf map16g;
  tdr::DevBuf tab_scaled;    // polar table x scale x res of a grid launch (uniform scale): staged here, mirrored in constant memory
  int map16g_log2 = -1;      // layout of map16g (-1: not built)
  int grid_phase_log2 = 0;   // x stride of the resident lattice of centres, if it is 2, 4 or 8 px (else 0)
  tdr::DevBuf scan_op;       // P_pad x N x 32 B
  tdr::DevBuf bin_counts, perm;
  tdr::DevBuf grid_key;      // 8 bytes: (min cost, first flat index) of the last tensor-core grid launch, packed
  bool grid_key_valid = false;
  int64_t perm_grid_n = -1;  // perm holds the binning of the RESIDENT grid centres (n of them); -1: it does not
  int score_impl = 0;        // 0 auto, 1 CUDA cores only, 2 tensor cores whenever usable
  int mma_tiles = 2;         // list kernel: 128-hypothesis tiles per CTA (tuning: TDR_MMA_TILES); two tiles share one
                             // streamed scan operand, 512 gather threads keep ~2000 records in flight per SM
  int mma_seg_shift = 2;     // log2 of the column-segment width of a bin (tuning: TDR_MMA_SEG_SHIFT)
  int mma_split = 2;         // gather threads per hypothesis row (tuning: TDR_MMA_SPLIT)
  int mma_a_tmem = 1;        // list kernel: gathered records to tensor memory instead of shared (tuning: TDR_MMA_A_TMEM)
  int mma_ring_cfg = 413;    // ring kernel: tiles * 10 + threads per row, + 100 operands in tensor memory, + 400 and 4-cell stages (tuning: TDR_MMA_RING_CFG)
  int mma_kernel = 0;        // 0 auto (integer kernel, else streamed-operand, else ring), 1 streamed-operand fp16 kernel only,
                             // 2 ring kernel only, 3 integer kernel first (= auto) (tuning: TDR_MMA_KERNEL)
  int mma_ctas = 0;          // cap on co-resident CTAs per SM (0 = as many as TMEM allows; tuning: TDR_MMA_CTAS)
  int mma_i8 = 1;            // integer 16-byte-record theta search (score_mma_i8.cu): 0 off, 1 where its error bound holds,
                             // 2 always (tests of the kernel itself; TDR_MMA_I8)
  int mma_tex = 0;           // integer kernel: every second cell through the texture pipe (TDR_MMA_TEX; measured slower: 5.35 against 4.89 ms)
  unsigned long long map8_tex = 0;     // cudaTextureObject_t over map8 (pitch-linear, border addressing)
  int mma_skip_rings = 1;    // integer kernel: do not gather lattice cells that meet no scan return under any candidate shift (TDR_MMA_SKIP_RINGS)
  int mma_i8_cfg = 141;      // integer kernel: tiles * 100 + gather threads per row * 10 + stages in flight per thread (TDR_MMA_I8_CFG)
  int mma_sort = 1;          // integer kernel, hypothesis order inside a super-tile: 0 = pixel row, 4-px segment (its records are
                             // row-major); 1 = Morton over 2 x 2-px cells with 4 x 2-px lines (TDR_MMA_SORT).  Equal while the
                             // kernel was balanced; 6 % faster (4.13 -> 3.87 ms) once the L1 data pipe became the bound
                             // (profiles/r02_sweep_i8.txt)
  tdr::DevBuf map8;          // rows*cols x 16 B: u16 fixed-point class distances (hi / lo bytes) + known, 4 x 2-px blocks
  bool map8_valid = false;
  float map8_q = 0.f;        // its quantum
  int map8_blocked = -1;     // its layout: 0 row-major, 1 blocks of 4 x 2 px
  // the previous scan's largest class-summed count, read back lazily: predicts which operand format the next scan fits
  // (u8 <= 255, fp16 <= 2048); the kernels re-check on the device
  int scan_max_seen = 0;
  bool scan_max_pending = false;
  cudaEvent_t scan_max_ev = nullptr;
  int* scan_max_pin = nullptr;
  int mma_grid_cap = 0;      // cap on the grid of the persistent score kernels (0 = none; TDR_MMA_GRID_CAP — tests use it to
                             // make every CTA walk many batches)
  int mma_st_shift = 10;     // log2 of the binning super-tile side in px (tuning: TDR_MMA_ST_SHIFT)

  // ---- polar table
  int n_theta = 0, n_r = 0;
  tdr::DevBuf tab;           // 2*P floats
  bool have_tab = false;
  uint64_t tab_version = 1;  // bumped by tdr_map_set_polar_table: the score kernels mirror the table in constant memory

  // ---- scan
  tdr::DevBuf pts;           // raw AoS copy
  int pts_stride = 0, pts_ioff = 0;
  int64_t n_pts = 0;
  tdr::DevBuf lut;           // int32[n_lut]
  int n_lut = 0, lut_classes = 0;
  tdr::DevBuf scan_img;      // C x P floats (col-major n_theta x n_r per class)
  int scan_theta = 0, scan_r = 0, scan_C = 0;
  bool have_scan = false;
  tdr::DevBuf scan_pack;     // P x 8 floats: per lattice cell (w_c*0.01*count_c for c<C, ..., slot7 = sum_c count_c)
  tdr::DevBuf hist;          // int32 bin counts
  // cfg5 batch binning in progress (tdr_refine_begin .. tdr_refine_counts / tdr_refine_rebuild_map)
  int refine_w = 0, refine_h = 0, refine_C = 0, refine_off_x = 0, refine_off_y = 0; float refine_res = 0.f;
  // host chunks: H2D on a copy stream into one of two staging slots while the previous chunk's kernel runs
  cudaStream_t copy_stream = nullptr; cudaEvent_t refine_copied[2] = {nullptr, nullptr}, refine_binned[2] = {nullptr, nullptr};
  tdr::DevBuf refine_stage[2]; int refine_slot = 0;

  // ---- filter
  tdr_filter_params fp{};
  bool have_params = false;
  std::vector<float> search_thetas;
  std::vector<int32_t> search_shifts;
  tdr::DevBuf d_search_thetas, d_search_shifts;
  tdr::Particles part[2];    // ping-pong (particles_ / new_particles_, particle_filter.cpp:187)
  int cur = 0;
  tdr::Particles ckpt;       // device-side snapshot of the particle set (tdr_pf_checkpoint / tdr_pf_restore)
  int64_t ckpt_uninit = 0;
  int64_t n_uninit = 0;      // particles still without a heading (have_init == false)
  // n_uninit is exact on the host except while a device recount is in flight: a gated particle keeps have_init = 0
  // through the theta search (state_particle.cpp:163-176 return before :205) and resampling multiplies or drops it,
  // so the set is recounted on the device (k_count_uninit -> pinned word, event) and read back lazily by the next
  // score call (tdr::sync_uninit) — no host round trip inside an update.
  bool uninit_pending = false;
  cudaEvent_t uninit_ev = nullptr;
  int* uninit_pin = nullptr;           // pinned host word the recount lands in
  tdr::DevBuf uninit_dev;              // its device-side counter
  // kernels whose dynamic shared-memory opt-in has been set ON THIS CONTEXT'S DEVICE (function attributes are per
  // device; a process may hold contexts on several)
  tdr::DevBuf ident_shifts; int ident_shifts_n = 0;   // 0 .. n_theta - 1 on the device (tracking passes of the integer kernel)
  // the particle-dependent preparation of a theta search (spatial sort, scale check) started on a side stream before the
  // scan is rasterised (score_i8_prepare_async); the search launch joins it
  cudaStream_t prep_stream = nullptr; cudaEvent_t prep_fork = nullptr, prep_done = nullptr; bool prep_pending = false;
  int count_scale = 1;                  // ranks sharing the particle set: kernel choices go by the GLOBAL count, so that
                                        // weights (hence resampled indices) do not depend on the number of ranks
  void* shard = nullptr;               // tdr::Shard (shard.cu): NCCL communicator, peer-mapped export slots
  uint64_t smem_optin = 0;
  uint64_t smem_optin_i8 = 0;          // the same for the variants of the integer score kernel
  uint64_t tab_id = 0;                 // process-unique id of the resident polar table (constant-memory mirrors key on it)
  tdr::DevBuf d_cw;          // class weights (16 floats)
  tdr::Particles all;        // multi-GPU: the all-gathered particle set (N = ranks * n_local) in global order
  tdr::DevBuf weights;       // raw -> normalised in place
  tdr::DevBuf raw_weights;   // copy of the raw weights of the last scoring (tdr_pf_keep_raw_weights)
  bool keep_raw = false;
  int64_t n_raw = 0;
  int64_t n_weights = 0;
  const float* ld_override = nullptr;  // device last_dist aligned with an all-gathered weight vector (multi-GPU)
  tdr::DevBuf prefix;        // running max of the order-exact prefix
  tdr::DevBuf seq_ws;        // workspace of the tiled order-exact accumulation
  int seq_impl = 0;          // 0 auto (tiled for long chains), 1 single-CTA kernel only (TDR_SEQ_IMPL)
  tdr::DevBuf idx;           // resampled indices
  tdr::DevBuf scal;          // small device scalars (sums, stats, argmax, pose)
  tdr::DevBuf pose_tmp;      // 4 x n floats (ml-state columns)
  tdr::PinBuf pin;           // pinned staging
  int64_t argmax = 0;
  bool have_argmax = false;

  // ---- stage timers (tdr_profile_enable): events bracket render / score / normalise / resample
  bool profiling = false;
  cudaEvent_t stage_ev[TDR_N_STAGES + 1] = {};
  bool stage_valid = false;

  // ---- grid (cfg4)
  tdr::DevBuf grid_centers, grid_costs, grid_shifts;
  int64_t grid_n = 0;
  int grid_shifts_n = 0;
  std::vector<int32_t> grid_shifts_host;
  // fused all-gather: peer-mapped full cost arrays (rank order), this rank's row offset
  float* grid_peers[TDR_MAX_PEERS] = {};
  int grid_n_peers = 0;
  int64_t grid_peer_row0 = 0;
  tdr::DevBuf grid_full;                 // this rank's full array (IPC-exported), followed by the exchange mailbox
  int64_t grid_full_floats = 0;          // floats of the cost array proper: the mailbox sits at the next 256-byte boundary
  uint64_t grid_epoch = 0;               // exchanges done (all ranks in lock step): picks the mailbox slot
  unsigned long long* grid_key_pin = nullptr;   // pinned host word the exchange kernel writes the reduced key to
  int edt_band = 0;                      // TDR_EDT_BAND: rows per column-sweep thread (default 64)
  int edt_impl = 0;                      // TDR_EDT_IMPL=1: scalar tap scan instead of the packed (DPX) row pass
  int grid_store_hint = 0;               // TDR_GRID_STORE_HINT=1: L2 evict-first policy on the cost stores
  bool grid_self_only = false;           // TDR_GRID_SELF_ONLY=1: store costs to the own array only (diagnostic)
  std::vector<void*> grid_opened;        // mappings to close
  float* grid_costs_ext = nullptr;       // caller-provided device buffer for the costs (tdr_grid_set_costs_buffer)
  int64_t grid_costs_ext_cap = 0;
};

namespace tdr {
// scalar slots in ctx->scal (floats unless noted)
enum {
  SC_SUM = 0, SC_NVALID, SC_MEAN, SC_BS, SC_NUNDER, SC_FALLBACK,     // stats (6 floats, ABI order)
  SC_S1, SC_S2, SC_REP, SC_ARGMAX /*int*/, SC_ARGVAL, SC_BSRAW /* sequential sum of lower squared deviations */,
  SC_MMA_MAXCOUNT = 12,     // int: max class-summed scan count (fp16 exactness check of the tensor-core path)
  SC_MMA_SCALE_RANGE = 14,  // 2 x uint32: min / max scale bits over the particles of an integer-kernel launch
  SC_MMA_BAILED = 13,       // int: the tensor-core kernel found counts above 2048 and left the work to the CUDA cores
  SC_CHAIN = 16,            // 8 chain totals
  SC_DBL = 32,              // doubles from here (8-byte aligned): sumsq, count_valid, count_under ...
  SC_POSE = 64,             // pose scalars
  SC_TOTAL = 256
};


inline void count_launch(tdr_ctx* c, int k = 1) { c->launches += k; }
// opt a kernel into more than 48 KB of dynamic shared memory once per context (= per device)
enum { OPTIN_SCORE_TRACK = 0, OPTIN_SCORE_SEARCH, OPTIN_SMALL_UPDATE, OPTIN_SCAN_BIN, OPTIN_EDT_ROWS, OPTIN_NORM_FUSED, OPTIN_GEO_POLAR,
       OPTIN_LIST_BASE = 16 /* + kernel variant */, OPTIN_RING_BASE = 32, OPTIN_TILE_BASE = 48 };
#define TDR_SMEM_OPTIN(ctx, bit, kernel, bytes)                                                                   \
  do {                                                                                                            \
    if (!((ctx)->smem_optin & (1ull << (bit)))) {                                                                 \
      TDR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));          \
      (ctx)->smem_optin |= 1ull << (bit);                                                                         \
    }                                                                                                             \
  } while (0)
uint64_t next_tab_id();
inline void stage_mark(tdr_ctx* c, int k) { if (c->profiling) cudaEventRecord(c->stage_ev[k], c->stream); }

// map_build.cu
int map_set_class_image(tdr_ctx*, const uint8_t*, int, int, int, const int32_t*, int, int, float);
int map_set_binary_layers(tdr_ctx*, const float*, int, int, int, float);
int map_set_dist_layers(tdr_ctx*, const float*, const uint8_t*, int, int, int, float);
int map_get_layers(tdr_ctx*, float*, uint8_t*);
int map_get_geo_layers(tdr_ctx*, float*);
int map_geo_resident(tdr_ctx*);          // builds ctx->geo_planar if stale
int local_geo_polar(tdr_ctx*, const float* dev_centers, int n, float scale, float res, float* dev_geo);
int active_pairwise(tdr_ctx*, const float* dev_maps, const int* dev_shifts, int n_cfg, int n_preds, float* dev_totals);
int map_from_seeds(tdr_ctx*, int rows, int cols, int C, float resolution);
int map_set_polygons(tdr_ctx*, const float* verts, const int32_t* poly_start, const int32_t* poly_class, int n_poly, int map_w, int map_h,
                     float rot, int C, float resolution, const int32_t* exclusive, int n_excl, float* layers_out);   // seedbits already filled on the device
// scan_render.cu
int scan_render(tdr_ctx*, bool polar, float res, float ang_res, int d0, int d1, float* dev_img_out);
int scan_pack(tdr_ctx*);
int scan_render_geometric(tdr_ctx*, bool polar, float res, float ang_res, int d0, int d1, int width, int height, float* dev_out);
int refine_begin(tdr_ctx*, float res, float cx, float cy, int width, int height, int C);
int refine_add(tdr_ctx*, const float* xy, const int32_t* cls, long long n, bool on_device);
int refine_counts(tdr_ctx*, uint8_t* maps_out);
int refine_rebuild_map(tdr_ctx*, float resolution);
int refine_bin(tdr_ctx*, const float* xy, const int32_t* cls, long long n, float res, float cx, float cy, int width,
               int height, int C, uint8_t* maps_out);
// score.cu
int score_particles(tdr_ctx*, float res);
int score_grid(tdr_ctx*, long long n, float scale, float res);
int local_polar(tdr_ctx*, const float* dev_centers, int n, float scale, float res, float* dev_dists, uint8_t* dev_mask);
int local_cart(tdr_ctx*, float cx, float cy, float rot, float res, int out_rows, int out_cols, float* dev_dists, uint8_t* dev_mask);
// score_mma.cu
int score_mma(tdr_ctx*, float res, bool grid_mode, long long n_items, float grid_scale, const int32_t* dev_shifts,
              const int32_t* host_shifts, int n_shifts, bool* used, bool track = false);
// score_mma_list.cu
int score_mma_list(tdr_ctx*, float res, bool grid_mode, long long n_items, float grid_scale, const int32_t* dev_shifts,
                   int n_shifts, bool* used);
// score_mma_i8.cu
int score_mma_i8(tdr_ctx*, float res, const int32_t* dev_shifts, int n_shifts, bool* used);
int score_mma_i8_track(tdr_ctx*, float res, bool* used);
int score_i8_prepare_async(tdr_ctx*);   // optional head start for the next theta search (pure search sets only)
int score_i8_prepare_join(tdr_ctx*);    // no-op when nothing is pending   // large tracked sets, in passes of 40 row shifts
// weights.cu
int normalize(tdr_ctx*, bool lazy_stddev = false);   // lazy: skip the lower-half deviation when no weight is NaN (stats[3] undefined then)
int build_prefix(tdr_ctx*);
int resample(tdr_ctx*, float u, long long M, long long i0, long long i1, Particles* src, Particles* dst);
int cache_ml_state(tdr_ctx*, const Particles& src);
int small_update(tdr_ctx*, float u, long long M, bool do_resample, bool* used);
int recount_uninit(tdr_ctx*, Particles& pt);   // asynchronous device recount of have_init == 0 over pt (see tdr_ctx::uninit_pending)
int sync_uninit(tdr_ctx*);                     // n_uninit exact on the host again
int exact_sums(tdr_ctx*, const float* const* cols, long long n, int ncols, float* totals_dev);
// pose.cu
int pose_of(tdr_ctx*, Particles& pt, float* mean, float* cov_mean, float* ml, float* cov_ml);
int grid_best(tdr_ctx*, const float* dev_costs, long long n, float* best_cost, long long* best_index);
int gmm_samples(tdr_ctx*, int num_samples, double* samples_host);
int propagate(tdr_ctx*, float tx, float ty, float omega, int scale_freeze, float pos_cov, float theta_cov, const float* z_dev, bool rng,
              unsigned long long seed, unsigned long long step, float* z_out_dev);
inline float* grid_costs_ptr(tdr_ctx* c) { return c->grid_costs_ext ? c->grid_costs_ext : c->grid_costs.as<float>(); }
}
