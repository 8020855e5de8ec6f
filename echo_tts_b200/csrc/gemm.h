// Host-visible description of one tcgen05 GEMM launch (see gemm_tc.cuh for the kernel).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace echo {

typedef __nv_bfloat16 bf16;

enum EpiMode : int { EPI_GENERIC = 0, EPI_SWIGLU = 1, EPI_QKV = 2,
                     EPI_ACCUM = 3 /* internal: lean instantiation of the pure residual accumulate, chosen by gemm_launch */,
                     EPI_RU = 4 /* fused DAC ResidualUnit: conv7 -> Snake -> conv1 -> + x, see GemmCall::B1 */,
                     EPI_RUW = 5 /* internal: EPI_RU with the conv weights resident in shared memory and the A rows of all
                                    taps loaded once per tile as one window; chosen by gemm_launch for 96 channels */,
                     EPI_CONV = 6 /* internal: lean form of the generic epilogue for the DAC convs -- bias (+ fp32 stream) ->
                                     fp32, Snake -> bf16 -- chosen by gemm_launch when nothing else is asked for */ };
enum ActMode : int { ACT_NONE = 0, ACT_GELU = 1, ACT_SNAKE = 2, ACT_TANH = 3, ACT_SIGMOID = 4, ACT_SILU = 5 };

struct QkvSection {
  bf16* out;            // [rows, sec_width] bf16
  const float* norm_w;  // [sec_width] per-(head,dim) RMSNorm weight, or null (no norm)
  int rope_heads;       // RoPE applied to 128-column groups with index < rope_heads (0 = none)
  int sigmoid;          // store sigmoid(v)
};

struct GemmParams {
  int M;        // rows per batch item
  int N;        // output columns
  int Kc;       // reduction length per tap
  int batches;  // batch items (A is addressed (k, row, batch))
  int taps;     // 1 for plain GEMM
  int tap_shift[8];  // row offset added to the A row coordinate for each tap
  int a_batch_div;   // A batch coordinate = batch / a_batch_div (>=1): lets several weight sets share one A
  int b_batch_rows;  // B row coordinate = batch * b_batch_rows + n0 (0: weights shared by all batch items)

  int epi;  // EpiMode (must match the template instantiation)
  // ---- EPI_GENERIC: v = (acc + bias[c]) * scale; v *= gate[g]; v += resid[r,c];
  //      out_f32[r,c] = v; out_bf16[r,c] = act(v).   c index for bias/alpha/gate is (col % col_mod).
  const float* bias;
  int bias_bstride;    // bias index += batch * bias_bstride
  float scale;
  const float* gate;   // fp32; index = (row / rows_per_gate) * gate_ld + (col % col_mod); rows_per_gate==0 -> row term 0
  int rows_per_gate;
  int gate_ld;
  const float* resid;  // fp32 [rows, ld_f32] (may alias out_f32: each element is read then written by one thread)
  float* out_f32;
  int ld_f32;
  bf16* out_bf16;
  int ld_bf16;
  int act;             // ActMode applied to the bf16 output only
  const float* alpha;  // snake alpha [col_mod]
  const float* alpha_inv;  // optional 1 / (alpha + 1e-9) [col_mod]
  int col_mod;
  float* part_out;     // EPI_ACCUM with split-K, set by gemm_launch from GemmCall::part_ws: slice sk stores its gated partial
  int64_t part_stride; // sum to part_out + sk * part_stride (same row / column layout as out_f32) instead of reducing into out_f32
  int atomic_out;     // EPI_GENERIC, resid == out_f32: add gate * acc into out_f32 with fp32 vector reductions instead of load-add-store
  int b_stream;       // set by gemm_launch: B is a large weight matrix read once per launch -> B loads L2 evict-first, A loads evict-last
  int no_b_prefetch;  // 1: B may have been written by the preceding kernel of the stream -> do not load it before the PDL wait
  int split_k;  // EPI_GENERIC with resid == out_f32 only: > 1 splits the K blocks over that many CTAs per tile (atomic adds)
  int n_valid;  // EPI_GENERIC: output columns >= n_valid are computed but not stored (N padded to a tile multiple); 0 = N
  // ---- EPI_SWIGLU: tile columns [0,BN/2) hold w1 rows, [BN/2,BN) the matching w3 rows;
  //      out_bf16[r, n0/2 + c] = silu(a) * b.     (uses out_bf16 / ld_bf16)
  // ---- EPI_QKV: N = nsec * sec_width, each 128-column group is normalised / rotated independently.
  QkvSection sec[4];
  int sec_width;
  const float* rope_cos;  // [pos, head_dim/2] fp32
  const float* rope_sin;
  int head_dim;    // 128 (DiT/encoders) or 64 (DAC post_module)
  int pos_period;  // position = pos_offset + pos_mult * (row % pos_period)
  int pos_offset;
  int pos_mult;
  float eps;
  // ---- EPI_RU (N == Kc == channels C, 96 or 192): the main loop is the dilated conv7 (taps x C -> C); its accumulator
  // gets bias + Snake(alpha) (the fields above), is re-quantised to bf16 INTO SHARED MEMORY as the A operand of a second
  // MMA with the 1 x 1 conv's weights (GemmCall::B1); that accumulator gets ru_bias1 + resid -> out_f32 (the fp32 stream)
  // and Snake(ru_alpha_out) -> out_bf16 (the next unit's conv input). One read + one write of the stream per unit.
  const float* ru_bias1;
  const float* ru_alpha_out;
  const float* ru_alpha_out_inv;
  // ---- EPI_QKV with 384-column tiles only, set by gemm_launch: which three 128-column groups (heads) column tile t
  // holds -- groups [3t], [3t+1] feed the N = 256 MMA (one per CTA of the pair, one per epilogue warp half), group [3t+2]
  // the N = 128 MMA (its rows split across the pair, its chunks across the halves). A value >= N / 128 is an empty slot.
  // The map spreads the expensive groups (RoPE + RMSNorm) one per tile into the shared slot instead of leaving them in
  // three consecutive tiles whose epilogue is then the kernel's tail (single wave, one TMEM stage).
  uint8_t tile_groups[96];
  int qkv_tiles;  // number of column tiles of the map (>= ceil(N / 384): tiles whose third slot is empty run one MMA per K step)
  // optional debug trace: 8 clock64 stamps per CTA (see gemm_tc.cuh); null in production
  long long* trace;
};

struct GemmCall {
  const bf16* A;           // [a_batches][M][lda] bf16, K contiguous
  int64_t lda;             // elements
  int64_t a_batch_stride;  // elements; 0 -> M*lda
  int64_t a_rows;          // rows of A per batch item the TMA map may touch; 0 -> M. Larger when A carries a halo of
                           // earlier rows in front of the M output rows (streaming convs: tap_shift is then >= 0)
  const bf16* B;           // [b_rows][ldb] bf16, K contiguous (nn.Linear weight layout)
  int64_t ldb;
  const bf16* B1;          // EPI_RU: the 1 x 1 conv's weights [C][ldb1] bf16
  int64_t ldb1;
  int64_t b_rows;  // 0 -> N (or N*batches when b_batch_rows is set)
  GemmParams p;
  int bn;  // tile N override (0 = auto)
  int cg;  // CTA-group override: 0 = auto, 1 = single CTA tiles, 2 = CTA pairs (tcgen05 cta_group::2)
  int split_k;  // 0 = auto, 1 = off, n = force n splits (ignored unless the epilogue is a pure residual accumulate)
  // Optional, residual accumulate only: a workspace of at least 4 planes of part_stride floats each (plane layout =
  // out_f32's). When gemm_launch splits K it then makes every slice store its partial sum to its own plane instead of
  // reducing into out_f32 with atomics, and reports the number of planes in *parts_used (0: out_f32 was updated as
  // usual). The caller must add the planes to out_f32 (rmsnorm_affine does it on the fly).
  float* part_ws;
  int64_t part_stride;
  int* parts_used;
};

cudaError_t gemm_launch(const GemmCall& c, cudaStream_t s);
int gemm_num_sms();
void gemm_set_deterministic(int on);  // 1: never use atomic split-K (bit-reproducible results)
int gemm_get_deterministic();

// Cached cuTensorMapEncodeTiled for a bf16 tensor whose dim0 is contiguous (rank 2 or 3; strides of dims >= 1 in
// BYTES). `out` points at a 128-byte CUtensorMap; box = (box0, box1, 1); swizzle span = row_bytes (128 / 64 / 32).
bool tma_map_bf16(void* out, const void* ptr, int rank, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2,
                  uint32_t box0, uint32_t box1, int row_bytes);

}  // namespace echo
