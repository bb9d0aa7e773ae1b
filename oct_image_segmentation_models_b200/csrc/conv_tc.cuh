// tcgen05 implicit-GEMM convolution ("tap-shifted descriptors over one TMA halo tile").
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <vector>

namespace octseg {

constexpr int kTcMaxKSteps = 40;   // k-steps (K=16 each) per input-channel chunk
constexpr int kTcTileW = 8;        // output tile: 8 px wide x 16 rows = 128 GEMM rows
constexpr int kTcTileH = 16;
constexpr int kTcMaxClasses = 16;

// Kernel parameters (passed by value as __grid_constant__).
struct TcConvParams {
  int n, h, w;                 // GEMM-row grid (== input grid; the low-res grid for up-conv)
  int tiles_x, tiles_y, n_tiles_n, num_tiles;
  int cin_chunks;              // number of input-channel chunks
  int planes_per_chunk;        // 8-channel planes per chunk
  int ksteps;                  // MMA k-steps per chunk
  int bgroup;                  // k-steps per weight stage
  int mt_x, mt_y;              // M-tiles per super-tile (one TMA halo box, mt_x*mt_y accumulators)
  int mt_x_log2;               // mt_x is a power of two
  int shuffle_pairs;           // mode 1: both x-parities of a plane live in the same n-tile
  int b_resident;              // 1: whole packed weight image stays in smem
  int n_cols;                  // MMA N per n-tile (multiple of 16, <= 256)
  int cols_valid;              // total valid GEMM columns over all n-tiles
  int pad_y, pad_x;            // halo before (rows / px)
  int box_w, box_h;            // halo tile size in px / rows
  int a_stages, b_stages;
  uint32_t a_stage_bytes, b_stage_bytes;
  uint32_t a_off[kTcMaxKSteps];   // byte offset of the k-step's first 8-channel half
  uint32_t a_lbo[kTcMaxKSteps];   // byte distance to its second half
  uint32_t a_desc_lo[kTcMaxKSteps];   // (a_off >> 4) | ((a_lbo >> 4) << 16): low descriptor word minus the stage base
  // epilogue
  int mode;                    // 0: plain blocked store, 1: 2x2 pixel-shuffle store (up-conv),
                               // 2: fused 1x1-conv + softmax head (activation never stored)
  int cout;                    // channels per parity (mode 1) or total (mode 0)
  int relu;
  int fp16;                    // 16-bit storage type of activations / packed weights: 0 = bf16, 1 = fp16
  const float *scale, *shift;  // [cout]
  __nv_bfloat16 *out;          // plane 0 of image 0 of the destination view
  long long out_img_stride;    // elements
  int out_h, out_w;
  __nv_bfloat16 *pool_out;     // mode 0: also write the 2x2 max-pooled tensor (NULL = off)
  long long pool_img_stride;
  int s2d;                     // data gradient of the up-conv on the LOW-res grid: the A stage holds the four pixel parities of dz
                               // as separate sub-tiles, fetched through four stride-2 tensor maps (s2d_maps, device memory)
  uint32_t s2d_part_bytes;     // smem distance between parity sub-tiles (128-byte aligned); a_tx_bytes = bytes TMA delivers per stage
  uint32_t a_tx_bytes;
  const CUtensorMap *s2d_maps;
  int pool_sum;                // 1: pool_out receives the 2x2 SUM of the fp32 outputs and the full-resolution tensor is not
                               // stored at all (data gradient of the up-conv: adjoint of the nearest x2 up-sampling)
  const float *head_w, *head_b; // mode 2: [cout][K] and [K]
  int head_k;
  float *probs;                // mode 2 outputs (NHWC fp32 / u8), either may be NULL
  uint8_t *labels;
  const __nv_bfloat16 *wpack;  // packed weights [n_tile][chunk][kstep][2][n_cols][8]
  int row_mul;                 // image rows per GEMM row (2 in row-pair mode, else 1)
  int scale_mod;               // entries of scale / shift / head tables (real cout); table index = column % scale_mod
  int split;                   // fp32-accurate mode: fp16 (hi, lo') plane pairs in, [main 8 | corr 8] column groups, (hi, lo') planes out
  float scale_mul;             // split mode: 1 / (power-of-two weight pre-scale), folded into the epilogue scale table
  int *overflow;               // split mode: device word set when an activation leaves the fp16 range
  double *stats;               // training forward: [2 * scale_mod] per-channel sum(z), sum(z^2) accumulated by the epilogue (NULL = off)
  int pdl_late;                // signal dependent launch after the last tile request (else at entry)
  int static_weights;          // packed weights are not written by earlier kernels of the stream (inference)
  uint32_t stage_off;          // mode 3: byte offset (from the epilogue tables) of the per-warp store staging
  int *status;                 // device word: non-zero = pipeline timeout code
  long long *dbg;              // optional [8] cycle counters of block 0 (NULL = off)
};

// Host-side plan for one conv block at one (n,h,w).
struct TcPlan {
  bool valid = false;
  TcConvParams p{};
  CUtensorMap tmap{};
  size_t smem_bytes = 0;
  int grid = 0;
  void *dev_maps = nullptr;    // s2d plans: the four parity tensor maps in device memory (owned; tc_release_plan)
  std::vector<void *> retired; // map buffers of earlier shapes: kept (512 B each) so that a re-plan never rewrites an address the
                               // GPU may still hold a cached descriptor of, or that queued kernels still read
};
void tc_release_plan(TcPlan *plan);

// Describes how GEMM K and N map to taps / channels; shared by the weight packer and
// the kernel's A-descriptor table.
struct TcGeometry {
  int kh, kw, cin, cout, ups;          // conv block
  int pt, pl;                          // padding before (rows / px); Keras "same": (k-1)/2
  int dy_min, dy_max, dx_min, dx_max;  // low-res tap range
  int planes_per_chunk, cin_chunks, ksteps, n_cols, n_tiles_n, cols_valid, bgroup;
  int box_w, box_h;                    // halo px / rows added around a super-tile
  int rows2;                           // set by the caller: every GEMM row yields TWO vertically adjacent output pixels
                                       // (geometry built for the 4x3 banded filter of tc_rowpair_weights)
  int stem_groups;                     // set by the caller: GEMM row = 8 adjacent pixels, columns = [plane][pixel][8 ch]
  int s2d;                             // tc_make_geometry_s2d: stride-2 3x3 conv over the pixel parities of the input (see there)
  // fp32-accurate split mode (tc_make_geometry_split): cin / cout above are PHYSICAL (2x logical).  Every logical
  // 8-channel input plane is a pair of fp16 planes (hi, lo' = (a - hi) * 2^11) and every logical 8-column output
  // group is the 16 GEMM columns [main 8 | corr 8]; the epilogue computes main + corr * 2^-11.
  int split;
  int cin_l, cout_l;                   // logical channel counts == strides of the fp32 weight tensor
  float wscale;                        // power-of-two pre-scale applied to the weights before the fp16 split
  // per k-step, per half: tap (dy,dx relative to dy_min/dx_min) and plane within chunk; tap -1 = zero
  int half_ty[kTcMaxKSteps][2], half_tx[kTcMaxKSteps][2], half_pl[kTcMaxKSteps][2];
};

bool tc_supported(int kh, int kw, int cin, int cout, int ups, int h, int w);
// pad_top/pad_left < 0: Keras "same" padding ((k-1)/2 before); otherwise explicit (data-gradient convs)
int tc_make_geometry(int kh, int kw, int cin, int cout, int ups, TcGeometry *g, int pad_top = -1, int pad_left = -1);
// Data gradient of the decoder's up-conv (x2 nearest up-sampling + 2x2 conv, reference models/unet.py:41-44) computed on
// the LOW-res grid: d(a)[Y][X] = sum_{a,b in {-1,0,1}} Weff[a][b]^T dz[2Y+a][2X+b], Weff[a][b] = sum of the forward taps
// (ky, kx) with ky in S(a), kx in S(b), S(-1) = {1}, S(0) = {0,1}, S(1) = {0}.  A stride-2 read is not a UMMA operand (rows
// of a core matrix are 16 B apart), so the four pixel parities of dz are fetched as four sub-tiles through stride-2 tensor
// maps; tap (a, b) is then parity (a != 0, b != 0) at box offset (a >= 0, b >= 0).  cz = channels of dz (<= 32), cout = the
// up-conv's input channels.  The packer reads the FORWARD kernel [2][2][cout][cz].
int tc_make_geometry_s2d(int cz, int cout, TcGeometry *g);
// fp32-accurate variant: error-compensated fp16 pairs on both operands (see TcGeometry::split); cin / cout logical
int tc_make_geometry_split(int kh, int kw, int cin, int cout, int ups, TcGeometry *g, int pad_top = -1, int pad_left = -1);
// power-of-two scale that brings max |w| of a layer into [2^3, 2^4) (fp16 keeps 11 bits for everything within 2^-17 of it)
float tc_split_weight_scale(const float *w, size_t count);
// Row-pair filter: GEMM row = pixel (2r, x) computes outputs (2r, x) and (2r+1, x) from the 4x3 input window
// rows 2r-1 .. 2r+2; column par*cout + co holds output row 2r+par.  Weff[a][b][ci][par*cout+co] = w[a-par][b][ci][co]
// for 0 <= a-par <= 2, else 0 (fp32 HWIO in, [4][3][cin][2*cout] out).  Geometry: tc_make_geometry(4, 3, cin,
// 2*cout, 0, &g, 1, 1) and g.rows2 = 1.  Nine taps per pixel become six: the A operand (the smem-bandwidth-bound
// side of narrow layers, 4 KB per K=16 step) is read 12 times per TWO output rows instead of 9 times per row.
void tc_rowpair_weights(const float *w_hwio, int cin, int cout, std::vector<float> *out);
// Banded 3x3 filter over pixel GROUPS for the tensor-core stem (see the definition): [3][3][1][cout] in,
// [3][3][8][8*cout] out; geometry tc_make_geometry(3, 3, 8, 8*cout, 0, &g) with g.stem_groups = 1.
void tc_stem_group_weights(const float *w, int cout, std::vector<float> *out);
// packs fp32 HWIO weights into the bf16 smem image the kernel streams (host memory)
void tc_pack_weights(const TcGeometry &g, const float *w_hwio, std::vector<uint16_t> *out, int fp16 = 0);
// same packing on the device, from fp32 weights in device memory (training)
int tc_pack_weights_device(const TcGeometry &g, const float *w_dev, int transposed, __nv_bfloat16 *out,
                           cudaStream_t st);
struct TcPackJob {
  TcGeometry g;
  const float *w;
  __nv_bfloat16 *out;
  long long total;
  int transposed;
};
int tc_pack_all_device(const TcPackJob *jobs_dev, int n_jobs, cudaStream_t st);
// pure host part of the plan (no CUDA driver needed): tiling, stage sizes, A-descriptor table
int tc_fill_params(const TcGeometry &g, int n, int h, int w, TcConvParams *p, size_t *smem_bytes);
struct TcEpilogue {
  int relu = 1;
  int fp16 = 0;
  const float *scale = nullptr, *shift = nullptr;
  View<__nv_bfloat16> out{};                 // unused when the head is fused
  __nv_bfloat16 *pool_out = nullptr;         // fused 2x2 max-pool (encoder-final blocks)
  long long pool_img_stride = 0;
  int pool_sum = 0;                          // pool_out = 2x2 sum instead of max, `out` is not written (out.ptr may be null)
  const float *head_w = nullptr, *head_b = nullptr;   // fused 1x1 conv + softmax head
  int head_k = 0;
  int static_weights = 0;                    // weights final before the stream's preceding kernels ran (inference)
  int *overflow = nullptr;                   // split mode: fp16 range overflow flag (device)
  double *stats = nullptr;                   // training forward: per-channel sum / sum of squares of the output (device, zeroed by the caller)
  float *probs = nullptr;
  uint8_t *labels = nullptr;
};
int tc_make_plan(const TcGeometry &g, const __nv_bfloat16 *in, int n, int h, int w,
                 const __nv_bfloat16 *wpack_dev, const TcEpilogue &epi, int *status_dev, TcPlan *plan);
int tc_launch(const TcPlan &plan, cudaStream_t st);
int tc_encode_map_4d(const void *base, int w, int h, int planes, int n, int box_w, int box_h, int box_planes,
                     CUtensorMap *out);
bool tc_head_fusable(int num_classes);

}  // namespace octseg
