// DSG-FD with the shipped criterion: KnowledgeDistillationKLDivLoss(T=2, reduction='sum') applied to
// [C,H,W] maps, i.e. softmax over H (kd_loss.py:28-34 with dim=1 == H), target = student*mask
// (detached), pred = teacher*mask.  Reference: gfl_deformable_detr_head_il.py:707-718 + kd_loss.py:12-43.
//
// Algorithmic bytes: read S + read T = 8 B per element (45.51 MB per 800x1333 image); no gradient reaches the
// student features (the target is detached): the only gradient is d loss / d mask rows.
//
// ONE pass over the features.  With x_h = S_h m_h / T (target logits), y_h = T_h m_h / T (pred logits):
//   KL(column) = sum_h q_h (x_h - y_h) - (lse_x - lse_y),     q = softmax(x), p = softmax(y)
//   d loss / d m_j = gcoef * sum_{h in run of box j} T_h (p_h - q_h)
//                  = gcoef * ( A_j / sum_y - B_j / sum_x ),   A_j = sum_run T_h e^{y_h},  B_j = sum_run T_h e^{x_h}
// so a column needs sum e^x, sum e^y, sum e^x (x - y) and, per run of rows with one owning box, (A_j, B_j): all of them
// are running sums down the rows, and the softmax normalisers are applied once at the bottom of the column.  The
// exponentials are taken WITHOUT subtracting the column maximum (cells outside boxes have logit 0, so the sums sit near
// H).  A (tile, channel) whose sums leave [2^-90, 2^100] or turn non-finite, and a tile with more runs than the record
// pools of its CTA hold, is not finished by the streaming kernel: every block leaves a bit mask of such channels, which a
// second, tiny launch works off with a plain three-sweep evaluation with exact maxima (the way the reference does).  Keeping that
// path in its own kernel keeps its registers and calls out of the streaming loop.  Cells outside every box are never read.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace dskd {

constexpr int kKlWarps = 8;       // warps per CTA of the redo kernel, and of the streaming kernel for large batches
constexpr int kKlMaxWarps = 16;   // streaming kernel: (channels per CTA / channels per pass) x row parts, at most
constexpr int kKlMaxSplit = 4;    // row parts of a column at most
constexpr int kKlPre = 16;        // owner rows a lane fetches at once in the prologue
// Measured on B200 (tools/kl_perf.py --tune, 16 images of 800x1333): two channels per pass, 4-row blocks loaded and consumed
// in place (no register ring), 4 CTAs x 8 warps per SM at 64 registers: 139 us; with a 2-stage ring at 127 registers
// (2 CTAs per SM) 168 us; one channel per pass (5-row blocks, 2 stages, 4 CTAs per SM) 169 us.
constexpr int kKlBlk = 4;         // rows per block: the unit of loading and of skipping rows without boxes (gradient kernels)
constexpr int kKlBlkFwd = 5;      //   ... forward-only kernels
constexpr int kKlChunk = 16;      // channels per CTA
constexpr int kKlSmallBatch = 4;  // images up to which the gradient kernel runs two row parts per column
constexpr int kKlPool = 192;      // (owner, lane, A, B) records a warp can park per pass before the normalisers are known
constexpr int kKlMaxH = 1600;     // rows per level (shared-memory tables: 136 B per row)
constexpr int kKlRedoCtas = 148;  // grid of the redo launch
constexpr float kLn2 = 0.6931471805599453f;

struct KlParams {
  DskdLevel levels[DSKD_MAX_LEVELS];
  const float* student[DSKD_MAX_LEVELS];
  const float* teacher[DSKD_MAX_LEVELS];
  float scale[DSKD_MAX_LEVELS];
  int block_start[DSKD_MAX_LEVELS + 1];
  int wtiles[DSKD_MAX_LEVELS];   // column tiles per level
  int wpt[DSKD_MAX_LEVELS];      // columns per tile (<= 32, balanced: ceil(W / wtiles))
  int num_levels, N, C;
  int chunk;                     // channels per CTA
  int max_blocks;                // row blocks of the tallest level (sizes the shared-memory tables)
  int max_h;
  int pool_cap;                  // (owner, lane, A, B) records a warp can park per channel
  int rsplit;                    // warps sharing one column (row parts)
  int dbg;                       // tuning experiments only (DSKD_KL_TUNE): 1 = no red.global at the bottom, 2 = no run ends
  float temperature, kscale;     // kscale = log2(e) / T: logits are kept in units of log 2
  int64_t cells_per_image;
  const int* owner;
  const float* cell_weight;      // per-cell mask weights instead of owners (CELL kernels)
  const float* rows;
  float* grad_rows;
  double* loss;
  unsigned* redo_mask;           // [blocks of the streaming grid] channels of the block's chunk left to the redo launch
  int num_blocks;
};

__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// KL of one column from its three sums (units of log 2 for ws): ln2 * ws / ss - ln(ss / st).  Both terms are first
// order, their difference second order: near ss == st the logarithm is taken as log1p of the relative difference.
__device__ __forceinline__ float kl_col(float ss, float st, float ws) {
  const float rs = __fdividef(1.f, ss), rt = __fdividef(1.f, st);
  const float d = (ss - st) * rt;  // (overflows for sums that are 2^128 apart: the plain logarithms then)
  const float lg = fabsf(d) < 0.5f ? log1pf(d) : logf(ss) - logf(st);
  return kLn2 * ws * rs - lg;
}
__device__ __forceinline__ bool kl_in_range(float ss, float st, float ws) {
  return ss > 0x1p-90f && ss < 0x1p100f && st > 0x1p-90f && st < 0x1p100f && fabsf(ws) < INFINITY;
}

// loads under a predicate (0 when off): cells outside boxes are never fetched
__device__ __forceinline__ float ld_stream_f1_ge0(const float* p, int key) {
  float v;
  asm("{\n\t.reg .pred q;\n\tsetp.ge.s32 q, %2, 0;\n\tmov.f32 %0, 0f00000000;\n\t"
      "@q ld.global.nc.L1::no_allocate.f32 %0, [%1];\n\t}"
      : "=f"(v)
      : "l"(p), "r"(key));
  return v;
}
__device__ __forceinline__ float ld_cached_f1_ge0(const float* p, int key) {
  float v;
  asm("{\n\t.reg .pred q;\n\tsetp.ge.s32 q, %2, 0;\n\tmov.f32 %0, 0f00000000;\n\t"
      "@q ld.global.nc.f32 %0, [%1];\n\t}"
      : "=f"(v)
      : "l"(p), "r"(key));
  return v;
}
__device__ __forceinline__ float ld_stream_f1_nz(const float* p, float key) {
  float v;
  asm("{\n\t.reg .pred q;\n\tsetp.neu.f32 q, %2, 0f00000000;\n\tmov.f32 %0, 0f00000000;\n\t"
      "@q ld.global.nc.L1::no_allocate.f32 %0, [%1];\n\t}"
      : "=f"(v)
      : "l"(p), "f"(key));
  return v;
}
// barrier of the warps that share channel slot `slot` (< 8); immediate ids, so that the CTA reserves 9 barriers, not 16
__device__ __forceinline__ void kl_slot_barrier(int slot, int threads) {
#define DSKD_BAR(ID) case ID - 1: asm volatile("bar.sync " #ID ", %0;" ::"r"(threads) : "memory"); break;
  switch (slot) { DSKD_BAR(1) DSKD_BAR(2) DSKD_BAR(3) DSKD_BAR(4) DSKD_BAR(5) DSKD_BAR(6) DSKD_BAR(7) DSKD_BAR(8) }
#undef DSKD_BAR
}
__device__ __forceinline__ void st_shared_b32(unsigned addr, int v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// Shared-memory tables of one column tile (all channels of the CTA share them).
struct KlTables {
  int* moff;          // [rows_padded][32] box masks: owner * C, -1 outside boxes / past the last row or column
  float* mw;          //                  cell masks: the cell's weight (aliases moff)
  unsigned* endm;     // [rows_padded] lanes whose run of one owner ends with this row
  unsigned* anym;     // [rows_padded] lanes with a box cell in this row
  unsigned short* act;// [nact] active row blocks (bit 15: some run ends inside the block)
};

// dynamic shared memory layout of the streaming kernel (the same arithmetic on both sides)
struct KlSmem {
  size_t endm, anym, pool, part, act, total;
  // rec_words: 32-bit words per run record: key + (A, B) per channel of the pass; nwarps: warps of the CTA;
  // part_words: partial sums per lane that the row parts of a column exchange (0: one warp per column)
  __host__ __device__ KlSmem(int max_blocks, int blk, bool grad, int pool_cap, int rec_words, int nwarps, int part_words) {
    const size_t rows = (size_t)max_blocks * blk;
    endm = rows * 32 * 4;
    anym = endm + rows * 4;
    pool = (anym + rows * 4 + 15) / 16 * 16;
    part = (pool + (grad ? (size_t)nwarps * pool_cap * rec_words * 4 : 0) + 15) / 16 * 16;
    act = part + (size_t)nwarps * 32 * part_words * 4;
    total = (act + (size_t)max_blocks * 2 + 15) / 16 * 16;
  }
};

// Which tile a block of the streaming grid owns (the redo launch decodes the same index).
struct KlTile {
  int lvl, img, chunk, w0, wn;
};
__device__ __forceinline__ KlTile kl_decode_tile(const KlParams& prm, int block) {
  KlTile t;
  t.lvl = 0;
#pragma unroll
  for (int k = 1; k < DSKD_MAX_LEVELS; ++k)
    if (k < prm.num_levels && block >= prm.block_start[k]) t.lvl = k;
  const int nchunks = (prm.C + prm.chunk - 1) / prm.chunk;
  int idx = block - prm.block_start[t.lvl];
  const int wt = idx % prm.wtiles[t.lvl];  // the column tiles of one plane are neighbours: they share its DRAM pages
  idx /= prm.wtiles[t.lvl];
  t.chunk = idx % nchunks;
  t.img = idx / nchunks;
  t.w0 = wt * prm.wpt[t.lvl];
  t.wn = min(prm.wpt[t.lvl], prm.levels[t.lvl].W - t.w0);
  return t;
}

// One streaming pass of a warp over the active row blocks of its channel plane.  GRAD: besides the three sums of the
// column, the (A, B) sums of every finished run are parked in the warp's record pool.
struct KlChan {
  const float* Sp;     // plane + w0 + lane
  const float* Tp;
  const float* rowsc;  // rows + c
  float ss, st, ws;    // sum e^x, sum e^y, sum e^x (x - y)     (x, y in units of log 2)
  int base;            // records parked so far (warp-uniform)
};

// NB = row blocks in a register ring (the loads of NB - 1 blocks in flight ahead of the arithmetic).  The kernel instantiates
// NB = 1 only: a block is loaded and consumed in place at 64 registers / 32 warps per SM; NB = 2 needed 127 registers and
// was slower at half the warps.
template <bool CELL, bool GRAD, int R, int NB>
__device__ __forceinline__ void kl_stream_pass(const KlTables& tb, KlChan& ch, const unsigned W, const int lane,
                                               const int nact, const float kscale, const unsigned pool_addr,
                                               const int pool_words, const int dbg) {
  struct Stage { float s[R], t[R]; };
  Stage stg[NB];
  float A = 0.f, B = 0.f;
  auto load = [&](Stage& sg, int entry) {
    const int r0 = (entry & 0x7fff) * R;
    unsigned o = (unsigned)r0 * W;
#pragma unroll
    for (int i = 0; i < R; ++i, o += W) {
      const int key = tb.moff[(r0 + i) * 32 + lane];
      if (CELL) {
        sg.s[i] = ld_stream_f1_nz(ch.Sp + o, __int_as_float(key));
        sg.t[i] = ld_stream_f1_nz(ch.Tp + o, __int_as_float(key));
      } else {
        sg.s[i] = ld_stream_f1_ge0(ch.Sp + o, key);
        sg.t[i] = ld_stream_f1_ge0(ch.Tp + o, key);
      }
    }
  };
  auto compute = [&](const Stage& sg, int entry, auto ends_tag) {
    constexpr bool ENDS = decltype(ends_tag)::value;  // some lane's run ends inside this block
    const int r0 = (entry & 0x7fff) * R;
    float m[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const int off = tb.moff[(r0 + i) * 32 + lane];
      m[i] = (CELL ? __int_as_float(off) : ld_cached_f1_ge0(ch.rowsc + (unsigned)off, off)) * kscale;
    }
    float pa[R], pb[R];  // ENDS: T_h e^y, T_h e^x of every row, so that the arithmetic of the block stays branch-free
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const float x = sg.s[i] * m[i], y = sg.t[i] * m[i];
      const float e = fast_ex2(x), f = fast_ex2(y);
      ch.ss += e;
      ch.st += f;
      ch.ws = fmaf(e, x - y, ch.ws);
      if (GRAD && !ENDS) {
        A = fmaf(sg.t[i], f, A);
        B = fmaf(sg.t[i], e, B);
      }
      if (ENDS) {
        pa[i] = sg.t[i] * f;
        pb[i] = sg.t[i] * e;
      }
    }
    if (ENDS) {
      unsigned em[R];
#pragma unroll
      for (int i = 0; i < R; ++i) em[i] = (dbg & 8) ? 0u : tb.endm[r0 + i];
#pragma unroll
      for (int i = 0; i < R; ++i) {
        A += pa[i];
        B += pb[i];
        if (em[i]) {  // warp-uniform: the run of some lane ends with this row
          // the lanes whose run ends here park (owner, lane, A, B) in consecutive pool slots
          const bool mine = (em[i] >> lane) & 1u;
          if (mine) {  // three planes of pool_words each: owner * C * 32 + lane, A, B
            const int slot = ch.base + __popc(em[i] & ((1u << lane) - 1u));
            st_shared_b32(pool_addr + 4u * slot, tb.moff[(r0 + i) * 32 + lane] * 32 + lane);
            st_shared_b32(pool_addr + 4u * (slot + pool_words), __float_as_int(A));
            st_shared_b32(pool_addr + 4u * (slot + 2 * pool_words), __float_as_int(B));
            A = 0.f;
            B = 0.f;
          }
          ch.base += __popc(em[i]);
        }
      }
    }
  };
  auto compute_any = [&](const Stage& sg, int entry) {
    if (GRAD && (entry & 0x8000) && !(dbg & 2)) compute(sg, entry, std::true_type{});
    else compute(sg, entry, std::false_type{});
  };
  // software pipeline over the active blocks: the loads of the next NB - 1 blocks are in flight during the arithmetic
#pragma unroll
  for (int j = 0; j < NB - 1; ++j)
    if (j < nact) load(stg[j], tb.act[j]);
  for (int k = 0; k < nact; k += NB) {
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      const int kk = k + u;
      if (kk < nact) {
        if (kk + NB - 1 < nact) load(stg[(u + NB - 1) % NB], tb.act[kk + NB - 1]);
        compute_any(stg[u], tb.act[kk]);
      }
    }
  }
}

// The same pass over TWO neighbouring channels (c, c + 1) at once.  The owner table, the predicates, the row offsets and
// the run-end bookkeeping are per column, not per channel, so they are paid once for both; the arithmetic runs on the
// packed fp32x2 pipe (FMUL2 / FADD2 / FFMA2: one issue slot per pair), every value a (channel c, channel c + 1) pair.
// The kernel is bound by instruction issue, not by DRAM (ncu: 75 % of the issue slots at 45 % of the DRAM roof with one
// channel per pass), which is what this halves.
struct KlChan2 {
  const float* Sp0;    // plane c + w0 + lane
  const float* Tp0;
  const float* Sp1;    // plane c + 1
  const float* Tp1;
  const float* rowsc;  // rows + c (8-byte aligned: c even, C even)
  float2 ss, st, ws;
  int base;
};

__device__ __forceinline__ void ld4_stream_ge0(float2& s, float2& t, const float* s0, const float* s1, const float* t0,
                                               const float* t1, int key) {
  asm("{\n\t.reg .pred q;\n\tsetp.ge.s32 q, %8, 0;\n\t"
      "mov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\tmov.f32 %2, 0f00000000;\n\tmov.f32 %3, 0f00000000;\n\t"
      "@q ld.global.nc.L1::no_allocate.f32 %0, [%4];\n\t@q ld.global.nc.L1::no_allocate.f32 %1, [%5];\n\t"
      "@q ld.global.nc.L1::no_allocate.f32 %2, [%6];\n\t@q ld.global.nc.L1::no_allocate.f32 %3, [%7];\n\t}"
      : "=f"(s.x), "=f"(s.y), "=f"(t.x), "=f"(t.y)
      : "l"(s0), "l"(s1), "l"(t0), "l"(t1), "r"(key));
}
__device__ __forceinline__ void ld4_stream_nz(float2& s, float2& t, const float* s0, const float* s1, const float* t0,
                                              const float* t1, float key) {
  asm("{\n\t.reg .pred q;\n\tsetp.neu.f32 q, %8, 0f00000000;\n\t"
      "mov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\tmov.f32 %2, 0f00000000;\n\tmov.f32 %3, 0f00000000;\n\t"
      "@q ld.global.nc.L1::no_allocate.f32 %0, [%4];\n\t@q ld.global.nc.L1::no_allocate.f32 %1, [%5];\n\t"
      "@q ld.global.nc.L1::no_allocate.f32 %2, [%6];\n\t@q ld.global.nc.L1::no_allocate.f32 %3, [%7];\n\t}"
      : "=f"(s.x), "=f"(s.y), "=f"(t.x), "=f"(t.y)
      : "l"(s0), "l"(s1), "l"(t0), "l"(t1), "f"(key));
}
__device__ __forceinline__ float2 ld_cached_f2_ge0(const float* p, int key) {
  float2 v;
  asm("{\n\t.reg .pred q;\n\tsetp.ge.s32 q, %3, 0;\n\tmov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\t"
      "@q ld.global.nc.v2.f32 {%0, %1}, [%2];\n\t}"
      : "=f"(v.x), "=f"(v.y)
      : "l"(p), "r"(key));
  return v;
}
__device__ __forceinline__ float2 ex2_2(float2 x) { return make_float2(fast_ex2(x.x), fast_ex2(x.y)); }

template <bool CELL, bool GRAD, int R, int NB>
__device__ __forceinline__ void kl_stream_pass2(const KlTables& tb, KlChan2& ch, const unsigned W, const int lane,
                                                const int nact, const float kscale, const unsigned pool_addr,
                                                const int pool_words, const int dbg) {
  struct Stage { float2 s[R], t[R]; };
  Stage stg[NB];
  float2 A = make_float2(0.f, 0.f), B = make_float2(0.f, 0.f);
  const float2 k2 = make_float2(kscale, kscale), neg1 = make_float2(-1.f, -1.f);
  auto load = [&](Stage& sg, int entry) {
    const int r0 = (entry & 0x7fff) * R;
    unsigned o = (unsigned)r0 * W;
#pragma unroll
    for (int i = 0; i < R; ++i, o += W) {
      const int key = tb.moff[(r0 + i) * 32 + lane];
      if (CELL) ld4_stream_nz(sg.s[i], sg.t[i], ch.Sp0 + o, ch.Sp1 + o, ch.Tp0 + o, ch.Tp1 + o, __int_as_float(key));
      else ld4_stream_ge0(sg.s[i], sg.t[i], ch.Sp0 + o, ch.Sp1 + o, ch.Tp0 + o, ch.Tp1 + o, key);
    }
  };
  auto compute = [&](const Stage& sg, int entry, auto ends_tag) {
    constexpr bool ENDS = decltype(ends_tag)::value;  // some lane's run ends inside this block
    const int r0 = (entry & 0x7fff) * R;
    float2 m[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const int off = tb.moff[(r0 + i) * 32 + lane];
      if (CELL) m[i] = make_float2(__int_as_float(off), __int_as_float(off));
      else m[i] = ld_cached_f2_ge0(ch.rowsc + (unsigned)off, off);
      m[i] = __fmul2_rn(m[i], k2);
    }
    float2 pa[R], pb[R];  // ENDS: T_h e^y, T_h e^x of every row, so that the arithmetic of the block stays branch-free
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const float2 x = __fmul2_rn(sg.s[i], m[i]), y = __fmul2_rn(sg.t[i], m[i]);
      const float2 e = ex2_2(x), f = ex2_2(y);
      ch.ss = __fadd2_rn(ch.ss, e);
      ch.st = __fadd2_rn(ch.st, f);
      ch.ws = __ffma2_rn(e, __ffma2_rn(y, neg1, x), ch.ws);
      if (GRAD && !ENDS) {
        A = __ffma2_rn(sg.t[i], f, A);
        B = __ffma2_rn(sg.t[i], e, B);
      }
      if (ENDS) {
        pa[i] = __fmul2_rn(sg.t[i], f);
        pb[i] = __fmul2_rn(sg.t[i], e);
      }
    }
    if (ENDS) {
      unsigned em[R];
#pragma unroll
      for (int i = 0; i < R; ++i) em[i] = (dbg & 8) ? 0u : tb.endm[r0 + i];
#pragma unroll
      for (int i = 0; i < R; ++i) {
        A = __fadd2_rn(A, pa[i]);
        B = __fadd2_rn(B, pb[i]);
        if (em[i]) {  // warp-uniform: the run of some lane ends with this row
          // the lanes whose run ends here park (owner, lane, A, B of both channels) in consecutive pool slots
          const bool mine = (em[i] >> lane) & 1u;
          if (mine) {  // five planes of pool_words each
            const unsigned a0 = pool_addr + 4u * (ch.base + __popc(em[i] & ((1u << lane) - 1u)));
            const unsigned pw = 4u * pool_words;
            st_shared_b32(a0, tb.moff[(r0 + i) * 32 + lane] * 32 + lane);
            st_shared_b32(a0 + pw, __float_as_int(A.x));
            st_shared_b32(a0 + 2 * pw, __float_as_int(B.x));
            st_shared_b32(a0 + 3 * pw, __float_as_int(A.y));
            st_shared_b32(a0 + 4 * pw, __float_as_int(B.y));
            A = make_float2(0.f, 0.f);
            B = make_float2(0.f, 0.f);
          }
          ch.base += __popc(em[i]);
        }
      }
    }
  };
  auto compute_any = [&](const Stage& sg, int entry) {
    if (GRAD && (entry & 0x8000) && !(dbg & 2)) compute(sg, entry, std::true_type{});
    else compute(sg, entry, std::false_type{});
  };
#pragma unroll
  for (int j = 0; j < NB - 1; ++j)
    if (j < nact) load(stg[j], tb.act[j]);
  for (int k = 0; k < nact; k += NB) {
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      const int kk = k + u;
      if (kk < nact) {
        if (kk + NB - 1 < nact) load(stg[(u + NB - 1) % NB], tb.act[kk + NB - 1]);
        compute_any(stg[u], tb.act[kk]);
      }
    }
  }
}

// Bottom of a column for the NC channels of a pass: the runs parked in the pool (plane 0: key, planes 1 + 2k / 2 + 2k:
// A / B of channel k) become d loss / d mask rows: neighbouring records of one owner are added up first (one segmented scan
// over the lanes for all channels: the segments are the same), then one red.global per segment and channel.
template <int NC>
__device__ __forceinline__ void kl_flush_pool(const int* pool, int pool_words, int count, int lane,
                                              const float (&rt)[NC] /* 1 / sum e^y of this lane's column */,
                                              const float (&rs)[NC] /* 1 / sum e^x */, float gcoef,
                                              float* __restrict__ growc, bool no_red) {
  constexpr unsigned kFull = 0xffffffffu;
  for (int j0 = 0; j0 < count; j0 += 32) {
    const int j = j0 + lane;
    const bool have = j < count;
    const int key = have ? pool[j] : -1 - lane;  // owner * C * 32 + lane of the column
    float g[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      // the normalisers of the record's column sit in the registers of its lane
      const float nt = __shfl_sync(kFull, rt[k], key & 31), ns = __shfl_sync(kFull, rs[k], key & 31);
      g[k] = 0.f;
      if (have)
        g[k] = gcoef * (__int_as_float(pool[j + (1 + 2 * k) * pool_words]) * nt -
                        __int_as_float(pool[j + (2 + 2 * k) * pool_words]) * ns);
    }
    const int own = key >> 5;
    const int prev = __shfl_up_sync(kFull, own, 1), next = __shfl_down_sync(kFull, own, 1);
    bool open = lane > 0 && prev == own;  // the segment continues to the left
    if (__any_sync(kFull, open)) {
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        float gv[NC];
#pragma unroll
        for (int k = 0; k < NC; ++k) gv[k] = __shfl_up_sync(kFull, g[k], d);
        const bool ov = __shfl_up_sync(kFull, (int)open, d);
        if (open && lane >= d) {
#pragma unroll
          for (int k = 0; k < NC; ++k) g[k] += gv[k];
          open = ov;
        }
      }
    }
    const bool tail = lane == 31 || next != own;
    if (have && tail && !no_red) {
      if constexpr (NC == 2) {  // (owner * C + c) is even and the rows are 8-byte aligned (pair_ok): one vector red
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(growc + own), "f"(g[0]), "f"(g[1]) : "memory");
      } else {
        atomicAdd(growc + own, g[0]);
      }
    }
  }
}

// Streaming kernel, [N,C,H,W] layout.  A CTA owns one column tile (level, image, <= 32 consecutive w) and a chunk of
// channels; each warp streams one channel plane at a time down ALL rows of the tile, lane = column, so every feature
// load is one coalesced row piece and the running sums of a column live in registers.  Once per CTA the owners of the
// tile's cells become a shared-memory table (owner * C per cell), the rows are cut into blocks of R and only the blocks
// that hold a box cell in some column are visited (the others add e^0 per row in closed form).  The (owner, A, B)
// records of finished runs wait in a per-warp pool until the bottom of the column.  The number of runs of a tile does
// not depend on the channel, so the CTA knows up front how many warps can run side by side: with more runs than one
// pool holds, 4 / 2 / 1 warps work with 2 / 4 / 8 pools each.  NC = 2: a warp takes two neighbouring channels per pass
// (kl_stream_pass2); NC = 1 is the scalar pass for an odd channel count.
template <bool CELL, bool GRAD, int NC, int R>
__global__ void __launch_bounds__(32 * kKlMaxWarps, 2) dsgfd_kl_stream_kernel(const __grid_constant__ KlParams prm) {
  extern __shared__ __align__(16) unsigned char kl_smem[];
  __shared__ double red[32];
  __shared__ int nact_s, nadd_s, need_s;
  __shared__ int part_ends[kKlMaxSplit];
  __shared__ unsigned redo_s;
  constexpr unsigned kFull = 0xffffffffu;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = (int)blockDim.x >> 5;
  const int rsplit = prm.rsplit;     // warps sharing one column: each takes a contiguous part of the active row blocks
  const int cw = nwarps / rsplit;    // channel slots: warps working on different channels side by side
  const KlTile tl = kl_decode_tile(prm, blockIdx.x);
  const int lvl = tl.lvl;
  const int H = prm.levels[lvl].H, C = prm.C;
  const unsigned W = (unsigned)prm.levels[lvl].W;
  const int nblk = (H + R - 1) / R, rows_pad = nblk * R;

  constexpr int kRec = 1 + 2 * NC;  // words per run record
  constexpr int kSums = 3 * NC;     // partial sums a row part hands over per lane
  const KlSmem lay(prm.max_blocks, R, GRAD, prm.pool_cap, kRec, nwarps, rsplit > 1 ? kSums : 0);
  KlTables tb;
  tb.moff = reinterpret_cast<int*>(kl_smem);
  tb.mw = reinterpret_cast<float*>(kl_smem);
  tb.endm = reinterpret_cast<unsigned*>(kl_smem + lay.endm);
  tb.anym = reinterpret_cast<unsigned*>(kl_smem + lay.anym);
  tb.act = reinterpret_cast<unsigned short*>(kl_smem + lay.act);
  float* part_sums = reinterpret_cast<float*>(kl_smem + lay.part);  // [nwarps][kSums][32]

  // ---- once per CTA: owner table, run ends, list of active row blocks, the row parts
  {
    const int64_t cell0 = (int64_t)tl.img * prm.cells_per_image + prm.levels[lvl].cell_offset + tl.w0 + lane;
    if (tid < kKlMaxSplit) part_ends[tid] = 0;
    if (tid == 0) redo_s = 0u;
    // A warp takes consecutive rows, lane = column: the owner table (owner * C per cell), and per row the lanes with a box
    // at all and the lanes whose run of one owner ends with this row.  Every row is fetched once: the next row's owners
    // are carried over (four rows in flight at a time).
    const int rpw = (rows_pad + nwarps - 1) / nwarps;
    const int h0 = warp * rpw, h1 = min(rows_pad, h0 + rpw);
    const bool col_in = lane < tl.wn;
    auto fetch = [&](int h) -> int {
      if (!(col_in && h < H)) return CELL ? 0 : -1;
      if (CELL) return __float_as_int(__ldg(prm.cell_weight + cell0 + (int64_t)h * W));
      return __ldg(prm.owner + cell0 + (int64_t)h * W);
    };
    // kKlPre rows in flight per lane: at 100 rows and 8 warps the whole share of a warp is one batch, one L2 latency
    int o = h0 < h1 ? fetch(h0) : -1;
    for (int h = h0; h < h1; h += kKlPre) {
      int nx[kKlPre];
#pragma unroll
      for (int u = 0; u < kKlPre; ++u) nx[u] = (h + u < h1) ? fetch(h + u + 1) : -1;
#pragma unroll
      for (int u = 0; u < kKlPre; ++u) {
        if (h + u < h1) {
          bool on, end = false;
          if (CELL) {
            on = __int_as_float(o) != 0.f;
            tb.moff[(h + u) * 32 + lane] = o;
          } else {
            on = o >= 0;
            end = on && nx[u] != o;
            tb.moff[(h + u) * 32 + lane] = on ? o * C : -1;
          }
          const unsigned am = __ballot_sync(kFull, on), em = __ballot_sync(kFull, end);
          if (lane == 0) { tb.anym[h + u] = am; tb.endm[h + u] = em; }
        }
        o = nx[u];
      }
    }
    __syncthreads();
    if (warp == 0) {
      int base = 0, nadd = 0;
      for (int b0 = 0; b0 < nblk; b0 += 32) {
        const int b = b0 + lane;
        bool on = false, ends = false;
        int valid = 0;
        if (b < nblk) {
          valid = min(R, H - b * R);
#pragma unroll
          for (int r = b * R; r < b * R + R; ++r) {
            on |= tb.anym[r] != 0u;
            ends |= tb.endm[r] != 0u;
          }
        }
        const unsigned bal = __ballot_sync(kFull, on);
        if (on) tb.act[base + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)(b | (ends ? 0x8000 : 0));
        base += __popc(bal);
        // rows of skipped blocks add e^0 to both sums; rows past H inside active blocks must not
        nadd += (b < nblk) ? (on ? valid - R : valid) : 0;
      }
      nadd = __reduce_add_sync(kFull, nadd);
      if (lane == 0) { nact_s = base; nadd_s = nadd; }
      if (GRAD) {
        // Row parts: part p streams the active blocks [p * q, (p + 1) * q).  A run that crosses into the next part is cut
        // at the part's last row (both pieces become records of the same owner and column: their sum is the run's).
        __syncwarp();
        const int q = (base + rsplit - 1) / rsplit;
        if (rsplit > 1 && lane >= 1 && lane < rsplit && lane * q < base) {
          const int k = lane * q - 1;  // last block of part lane - 1
          const int blk = tb.act[k] & 0x7fff, r = blk * R + R - 1;
          if (tb.anym[r] != 0u) {
            tb.endm[r] = tb.anym[r];
            tb.act[k] = (unsigned short)(blk | 0x8000);
          }
        }
        __syncwarp();
        for (int k = lane; k < base; k += 32) {  // records every part must be able to park
          const int r0 = (tb.act[k] & 0x7fff) * R;
          int e = 0;
#pragma unroll
          for (int r = r0; r < r0 + R; ++r) e += __popc(tb.endm[r]);
          if (e) atomicAdd(&part_ends[rsplit > 1 ? k / q : 0], e);
        }
        __syncwarp();
        int need = lane < kKlMaxSplit ? part_ends[lane] : 0;
#pragma unroll
        for (int d = 1; d < kKlMaxSplit; d <<= 1) need = max(need, __shfl_xor_sync(kFull, need, d));
        if (lane == 0) need_s = need;
      }
    }
    __syncthreads();
  }
  const int nact = nact_s;
  if (nact == 0) {  // no box touches this tile: every column's KL is exactly 0
    if (tid == 0) prm.redo_mask[blockIdx.x] = 0u;
    return;
  }
  const int c_begin = tl.chunk * prm.chunk, c_end = min(C, c_begin + prm.chunk);
  // channel slots working side by side: each warp needs pool room for every run of its row part
  int share = 1;
  if (GRAD) {
    const int need = need_s;
    while (share < cw && need > share * prm.pool_cap) share *= 2;
    if (need > share * prm.pool_cap) {  // not even one slot with every pool: the whole block is left to the redo launch
      if (tid == 0) prm.redo_mask[blockIdx.x] = 0xffffffffu >> (32 - (c_end - c_begin));
      return;
    }
  }
  const int part = warp / cw, cslot = warp - part * cw;
  const int wstep = cw / share;  // active channel slots: 0 .. wstep - 1; slot s of part p owns the pools p * cw + s * share ..
  // this warp's records: planes (key, A, B per channel) of share * pool_cap words each
  const int pool_words = share * prm.pool_cap;
  const int* pool = reinterpret_cast<const int*>(kl_smem + lay.pool) + (size_t)(part * cw + cslot * share) * kRec * prm.pool_cap;
  const unsigned pool_addr = (unsigned)__cvta_generic_to_shared(pool);
  double kl_total = 0.0;

  if (cslot < wstep) {
    const bool col_ok = lane < tl.wn;
    // this warp's part of the active blocks
    const int q = (nact + rsplit - 1) / rsplit;
    const int k0 = min(nact, part * q), k1 = min(nact, k0 + q);
    KlTables tbp = tb;
    tbp.act = tb.act + k0;
    const int nmine = k1 - k0;
    const float nadd = (float)nadd_s;
    for (int c = c_begin + NC * cslot; c < c_end; c += NC * wstep) {
      const int64_t plane = ((int64_t)tl.img * C + c) * ((int64_t)H * W) + tl.w0 + lane;
      float sums[kSums];  // per channel: sum e^x, sum e^y, sum e^x (x - y)
      int base;
      if constexpr (NC == 2) {
        KlChan2 ch;
        ch.Sp0 = prm.student[lvl] + plane;
        ch.Tp0 = prm.teacher[lvl] + plane;
        ch.Sp1 = ch.Sp0 + (int64_t)H * W;
        ch.Tp1 = ch.Tp0 + (int64_t)H * W;
        ch.rowsc = CELL ? nullptr : prm.rows + c;
        // keep the bases in registers: every address is then one IMAD.WIDE.U32 of a 32-bit element offset
        asm volatile("" : "+l"(ch.Sp0), "+l"(ch.Tp0), "+l"(ch.Sp1), "+l"(ch.Tp1), "+l"(ch.rowsc));
        ch.ss = ch.st = ch.ws = make_float2(0.f, 0.f);
        ch.base = 0;
        kl_stream_pass2<CELL, GRAD, R, 1>(tbp, ch, W, lane, nmine, prm.kscale, pool_addr, pool_words, prm.dbg);
        sums[0] = ch.ss.x; sums[1] = ch.st.x; sums[2] = ch.ws.x;
        sums[3] = ch.ss.y; sums[4] = ch.st.y; sums[5] = ch.ws.y;
        base = ch.base;
      } else {
        KlChan ch;
        ch.Sp = prm.student[lvl] + plane;
        ch.Tp = prm.teacher[lvl] + plane;
        ch.rowsc = CELL ? nullptr : prm.rows + c;
        asm volatile("" : "+l"(ch.Sp), "+l"(ch.Tp), "+l"(ch.rowsc));
        ch.ss = ch.st = ch.ws = 0.f;
        ch.base = 0;
        kl_stream_pass<CELL, GRAD, R, 1>(tbp, ch, W, lane, nmine, prm.kscale, pool_addr, pool_words, prm.dbg);
        sums[0] = ch.ss; sums[1] = ch.st; sums[2] = ch.ws;
        base = ch.base;
      }
      if (rsplit > 1) {
        // the row parts of this column meet: every part adds the partial sums up in the same order (identical totals)
        float* mine = part_sums + (size_t)warp * kSums * 32 + lane;
#pragma unroll
        for (int k = 0; k < kSums; ++k) mine[k * 32] = sums[k];
        kl_slot_barrier(cslot, 32 * rsplit);
#pragma unroll
        for (int k = 0; k < kSums; ++k) sums[k] = 0.f;
        for (int p = 0; p < rsplit; ++p) {
          const float* theirs = part_sums + (size_t)(p * cw + cslot) * kSums * 32 + lane;
#pragma unroll
          for (int k = 0; k < kSums; ++k) sums[k] += theirs[k * 32];
        }
        kl_slot_barrier(cslot, 32 * rsplit);  // the partial sums may be overwritten by the next pass
      }
      // ---- bottom of the column: KL, parked runs
      bool in_range = true;
      float kl = 0.f, rt[NC], rs[NC];
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        const float ss = sums[3 * k] + nadd, st = sums[3 * k + 1] + nadd, ws = sums[3 * k + 2];
        in_range = in_range && kl_in_range(ss, st, ws);
        kl += kl_col(ss, st, ws);
        rt[k] = __fdividef(1.f, st);
        rs[k] = __fdividef(1.f, ss);
      }
      const bool ok = __all_sync(kFull, !col_ok || in_range);
      if (ok) {
        if (col_ok && part == 0) kl_total += (double)kl;
        if (GRAD && !(prm.dbg & 4)) {
          const float gcoef = prm.scale[lvl] * prm.temperature / (float)H;  // d loss / d pred = scale * (T/H) * (p - q)
          __syncwarp();
          kl_flush_pool<NC>(pool, pool_words, base, lane, rt, rs, gcoef, prm.grad_rows + c, prm.dbg & 1);
          __syncwarp();
        }
      } else if (lane == 0 && part == 0) {
        // exponentials out of range for the unshifted sums: these (tile, channel)s are redone with exact maxima
        atomicOr(&redo_s, (NC == 2 ? 3u : 1u) << (c - c_begin));
      }
    }
  }
  // loss = scale * T^2 / H * sum over columns of sum_h q (log q - log p)
  double tot = block_sum(kl_total, red);
  if (tid == 0) {
    prm.redo_mask[blockIdx.x] = redo_s;
    if (tot != 0.0)
      atomicAdd(prm.loss, tot * (double)prm.scale[lvl] * (double)prm.temperature * (double)prm.temperature / (double)H);
  }
}

// The redo launch: (tile, channel) pairs the streaming kernel could not finish, evaluated the way the reference does:
// column maxima, sums against them, and a gradient sweep that issues one red.global per run.  A CTA takes one list entry
// at a time, a warp one channel, lane = column.  Rare (logits beyond +-60 / T, or hundreds of boxes crossing one tile),
// so plain code.
template <bool CELL, bool GRAD>
__global__ void __launch_bounds__(32 * kKlWarps) dsgfd_kl_redo_kernel(const __grid_constant__ KlParams prm) {
  extern __shared__ __align__(16) unsigned char kl_smem[];
  __shared__ double red[32];
  int* moff = reinterpret_cast<int*>(kl_smem);
  float* mw = reinterpret_cast<float*>(kl_smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float Temp = prm.temperature;
  // the CTA's share of the masks in one coalesced read: normally all zero, and the launch is over after one load
  const int per = (prm.num_blocks + (int)gridDim.x - 1) / (int)gridDim.x;
  const int b_first = blockIdx.x * per, b_last = min(prm.num_blocks, b_first + per);
  {
    unsigned any = 0u;
    for (int b = b_first + tid; b < b_last; b += 32 * kKlWarps) any |= prm.redo_mask[b];
    if (!__syncthreads_or(any != 0u)) return;
  }
  for (int b = b_first; b < b_last; ++b) {
    const unsigned todo = prm.redo_mask[b];
    if (todo == 0u) continue;  // uniform over the CTA
    const KlTile tl = kl_decode_tile(prm, b);
    const int lvl = tl.lvl;
    const int H = prm.levels[lvl].H, C = prm.C;
    const unsigned W = (unsigned)prm.levels[lvl].W;
    const int64_t cell0 = (int64_t)tl.img * prm.cells_per_image + prm.levels[lvl].cell_offset + tl.w0;
    __syncthreads();
    for (int i = tid; i < H * 32; i += 32 * kKlWarps) {
      const int h = i >> 5, l = i & 31;
      const bool in = l < tl.wn;
      if (CELL) mw[i] = in ? __ldg(prm.cell_weight + cell0 + (int64_t)h * W + l) : 0.f;
      else {
        const int o = in ? __ldg(prm.owner + cell0 + (int64_t)h * W + l) : -1;
        moff[i] = o >= 0 ? o * C : -1;
      }
    }
    __syncthreads();
    const int c_begin = tl.chunk * prm.chunk;
    const float gcoef = prm.scale[lvl] * Temp / (float)H;
    double kl_total = 0.0;
    // warp w takes the w-th, (w + 8)-th ... set bit of the mask
    int nth = 0;
    for (unsigned rest = todo; rest; rest &= rest - 1u, ++nth) {
      if ((nth % kKlWarps) != warp) continue;
      const int c = c_begin + (__ffs(rest) - 1);
      const int64_t plane = ((int64_t)tl.img * C + c) * ((int64_t)H * W) + tl.w0 + lane;
      const float* __restrict__ Sp = prm.student[lvl] + plane;
      const float* __restrict__ Tp = prm.teacher[lvl] + plane;
      const float* __restrict__ rowsc = CELL ? nullptr : prm.rows + c;
      auto logits = [&](int h, float& x, float& y, float& tf, int& off) {
        x = 0.f; y = 0.f; tf = 0.f;
        float m = 0.f;
        if (CELL) {
          m = mw[h * 32 + lane];
          off = m != 0.f ? 0 : -1;
        } else {
          off = moff[h * 32 + lane];
          if (off >= 0) m = rowsc[off];
        }
        if (off >= 0) {
          tf = Tp[h * W];
          x = __fdiv_rn(Sp[h * W] * m, Temp);
          y = __fdiv_rn(tf * m, Temp);
        }
      };
      float Ms = -INFINITY, Mt = -INFINITY;
      for (int h = 0; h < H; ++h) {
        float x, y, tf; int off;
        logits(h, x, y, tf, off);
        Ms = fmaxf(Ms, x);
        Mt = fmaxf(Mt, y);
      }
      float ss = 0.f, st = 0.f, ws = 0.f;
      for (int h = 0; h < H; ++h) {
        float x, y, tf; int off;
        logits(h, x, y, tf, off);
        const float ex = expf(x - Ms);
        ss += ex;
        st += expf(y - Mt);
        ws = fmaf(ex, x - y, ws);
      }
      if (lane < tl.wn)
        kl_total += (double)ws / (double)ss - (((double)Ms - (double)Mt) + log((double)ss / (double)st));
      if (GRAD) {
        float* __restrict__ growc = prm.grad_rows + c;
        float acc = 0.f;
        for (int h = 0; h < H; ++h) {
          float x, y, tf; int off;
          logits(h, x, y, tf, off);
          if (off < 0) continue;
          acc = fmaf(tf, expf(y - Mt) / st - expf(x - Ms) / ss, acc);
          const int nx = (h + 1 < H) ? moff[(h + 1) * 32 + lane] : -1;
          if (nx != off) {
            atomicAdd(growc + off, gcoef * acc);
            acc = 0.f;
          }
        }
      }
    }
    double tot = block_sum(kl_total, red);
    if (tid == 0 && tot != 0.0)
      atomicAdd(prm.loss, tot * (double)prm.scale[lvl] * (double)Temp * (double)Temp / (double)H);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Encoder-memory layout [S, N, C] (channel fastest; head_il.py:866-880 takes [N,C,H,W] *views* of it for sg_out /
// fg_only): per-cell mask weights only, forward only (those masks are constants and the target is detached, so nothing
// receives a gradient: SURVEY A3-var).  A warp owns one column (level, image, w) and 128 channels, lane = 4 consecutive
// channels: every feature load is one aligned 512 B piece of a token row, the weight of a cell is warp-uniform, and
// cells with weight 0 are skipped exactly (they add e^0 to both sums in closed form).
constexpr int kSncWarps = 8;
constexpr int kSncRows = 4;  // rows loaded ahead of the arithmetic (8 x 512 B in flight per warp)

struct KlSncParams {
  DskdLevel levels[DSKD_MAX_LEVELS];
  float scale[DSKD_MAX_LEVELS];
  int task_start[DSKD_MAX_LEVELS + 1];  // first warp task of each level; task = ((img * W + w) * groups + group)
  int num_levels, N, C, groups;         // groups of 128 channels
  float temperature, kscale;
  int64_t cells_per_image;
  const float* student;
  const float* teacher;
  const float* cell_weight;
  double* loss;
  unsigned* redo_mask;                  // [tasks / groups] bit g: group g of the column is left to the redo launch
};

struct KlSncTask {
  int lvl, img, w, group;
};
__device__ __forceinline__ KlSncTask kl_snc_decode(const KlSncParams& prm, int task) {
  KlSncTask t;
  t.lvl = 0;
#pragma unroll
  for (int k = 1; k < DSKD_MAX_LEVELS; ++k)
    if (k < prm.num_levels && task >= prm.task_start[k]) t.lvl = k;
  int idx = task - prm.task_start[t.lvl];
  t.group = idx % prm.groups;
  idx /= prm.groups;
  t.w = idx % prm.levels[t.lvl].W;
  t.img = idx / prm.levels[t.lvl].W;
  return t;
}

__global__ void __launch_bounds__(32 * kSncWarps, 4) dsgfd_kl_snc_kernel(const __grid_constant__ KlSncParams prm) {
  __shared__ double red[32];
  constexpr unsigned kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int task = blockIdx.x * kSncWarps + (threadIdx.x >> 5);
  double kl_total = 0.0;
  float loss_scale = 0.f;
  if (task < prm.task_start[prm.num_levels]) {
    const KlSncTask tk = kl_snc_decode(prm, task);
    const int H = prm.levels[tk.lvl].H, W = prm.levels[tk.lvl].W, C = prm.C;
    const int c0 = tk.group * 128 + lane * 4;
    const bool ch_ok = c0 < C;  // C % 4 == 0 (host): a lane's four channels exist together
    const float* __restrict__ wcol = prm.cell_weight + (int64_t)tk.img * prm.cells_per_image + prm.levels[tk.lvl].cell_offset + tk.w;
    // token (h, w) of image img: ((cell_offset + h * W + w) * N + img) * C
    const int64_t tok0 = ((prm.levels[tk.lvl].cell_offset + tk.w) * (int64_t)prm.N + tk.img) * C + (ch_ok ? c0 : 0);
    const int64_t pitch = (int64_t)W * prm.N * C;
    const float* __restrict__ Sp = prm.student + tok0;
    const float* __restrict__ Tp = prm.teacher + tok0;
    float ss[4] = {0.f, 0.f, 0.f, 0.f}, st[4] = {0.f, 0.f, 0.f, 0.f}, ws[4] = {0.f, 0.f, 0.f, 0.f};
    int visited = 0;
    for (int h0 = 0; h0 < H; h0 += 32) {
      // lane = row: the weights of 32 rows of this column, then the rows with a weight one after the other
      const float wv = (h0 + lane < H) ? __ldg(wcol + (int64_t)(h0 + lane) * W) : 0.f;
      unsigned todo = __ballot_sync(kFull, wv != 0.f);
      visited += __popc(todo);
      while (todo) {
        float4 sv[kSncRows], tv[kSncRows];
        float mk[kSncRows];
        // slots past the last row of this batch compute e^0 like a cell of weight 0: counted as visited so that the
        // closed-form term below takes them out again
        visited += max(0, kSncRows - __popc(todo));
#pragma unroll
        for (int i = 0; i < kSncRows; ++i) {
          const int r = todo ? __ffs(todo) - 1 : -1;
          todo &= todo - 1u;  // (0 stays 0)
          mk[i] = r >= 0 ? __shfl_sync(kFull, wv, r) * prm.kscale : 0.f;
          if (r >= 0 && ch_ok) {
            sv[i] = ld_stream_f4(Sp + (h0 + r) * pitch);
            tv[i] = ld_stream_f4(Tp + (h0 + r) * pitch);
          } else {
            sv[i] = tv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int i = 0; i < kSncRows; ++i) {
          const float xs[4] = {sv[i].x * mk[i], sv[i].y * mk[i], sv[i].z * mk[i], sv[i].w * mk[i]};
          const float ys[4] = {tv[i].x * mk[i], tv[i].y * mk[i], tv[i].z * mk[i], tv[i].w * mk[i]};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float e = fast_ex2(xs[j]), f = fast_ex2(ys[j]);
            ss[j] += e;
            st[j] += f;
            ws[j] = fmaf(e, xs[j] - ys[j], ws[j]);
          }
        }
      }
    }
    // cells never visited have logit 0: e^0 each
    const float nadd = (float)(H - visited);
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ss[j] += nadd;
      st[j] += nadd;
      ok = ok && kl_in_range(ss[j], st[j], ws[j]);
    }
    if (__all_sync(kFull, !ch_ok || ok)) {
      if (ch_ok) {
        float k4 = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) k4 += kl_col(ss[j], st[j], ws[j]);
        kl_total = (double)k4;
      }
    } else if (lane == 0) {
      atomicOr(prm.redo_mask + task / prm.groups, 1u << tk.group);
    }
    loss_scale = prm.scale[tk.lvl] * prm.temperature * prm.temperature / (float)H;
  }
  // the warps of a CTA may sit on different levels: scale per warp before the block sum
  double tot = block_sum(kl_total * (double)loss_scale, red);
  if (threadIdx.x == 0 && tot != 0.0) atomicAdd(prm.loss, tot);
}

// redo launch of the [S,N,C] kernel: exact maxima, lane = channel, one warp per flagged (column, 128-channel group)
__global__ void __launch_bounds__(32 * kSncWarps) dsgfd_kl_snc_redo_kernel(const __grid_constant__ KlSncParams prm) {
  __shared__ double red[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int columns = prm.task_start[prm.num_levels] / prm.groups;
  const float Temp = prm.temperature;
  for (int col0 = blockIdx.x * kSncWarps; col0 < columns; col0 += gridDim.x * kSncWarps) {
    const int col = col0 + warp;
    const unsigned todo = col < columns ? prm.redo_mask[col] : 0u;
    double kl_total = 0.0;
    for (unsigned rest = todo; rest; rest &= rest - 1u) {
      const int group = __ffs(rest) - 1;
      const KlSncTask tk = kl_snc_decode(prm, col * prm.groups + group);
      const int H = prm.levels[tk.lvl].H, W = prm.levels[tk.lvl].W, C = prm.C;
      const float* __restrict__ wcol = prm.cell_weight + (int64_t)tk.img * prm.cells_per_image + prm.levels[tk.lvl].cell_offset + tk.w;
      const int64_t pitch = (int64_t)W * prm.N * C;
      double kl_group = 0.0;
      for (int cc = lane; cc < 128; cc += 32) {
        const int c = group * 128 + cc;
        if (c >= C) break;
        const int64_t tok0 = ((prm.levels[tk.lvl].cell_offset + tk.w) * (int64_t)prm.N + tk.img) * C + c;
        const float* __restrict__ Sp = prm.student + tok0;
        const float* __restrict__ Tp = prm.teacher + tok0;
        auto logits = [&](int h, float& x, float& y) {
          const float wv = wcol[(int64_t)h * W];
          x = wv != 0.f ? __fdiv_rn(Sp[h * pitch] * wv, Temp) : 0.f;
          y = wv != 0.f ? __fdiv_rn(Tp[h * pitch] * wv, Temp) : 0.f;
        };
        float Ms = -INFINITY, Mt = -INFINITY;
        for (int h = 0; h < H; ++h) {
          float x, y;
          logits(h, x, y);
          Ms = fmaxf(Ms, x);
          Mt = fmaxf(Mt, y);
        }
        float ss = 0.f, st = 0.f, ws = 0.f;
        for (int h = 0; h < H; ++h) {
          float x, y;
          logits(h, x, y);
          const float ex = expf(x - Ms);
          ss += ex;
          st += expf(y - Mt);
          ws = fmaf(ex, x - y, ws);
        }
        kl_group += (double)ws / (double)ss - (((double)Ms - (double)Mt) + log((double)ss / (double)st));
      }
      kl_total += kl_group * (double)prm.scale[tk.lvl] * (double)Temp * (double)Temp / (double)H;
    }
    double tot = block_sum(kl_total, red);
    if (threadIdx.x == 0 && tot != 0.0) atomicAdd(prm.loss, tot);
  }
}

}  // namespace dskd

using namespace dskd;

// one 32-bit channel mask per block of the streaming grid, for the smallest chunk (1 channel per block)
extern "C" int64_t dskd_dsgfd_kl_workspace_bytes(int32_t N, int32_t num_levels, const DskdLevel* levels, int32_t C) {
  if (N < 0 || num_levels <= 0 || num_levels > DSKD_MAX_LEVELS || levels == nullptr || C <= 0) return -1;
  int64_t tiles = 0, columns = 0;
  for (int l = 0; l < num_levels; ++l) {
    tiles += (levels[l].W + 31) / 32;
    columns += levels[l].W;
  }
  // NCHW: a word per block of the streaming grid; [S,N,C]: a word per (image, level, w) column
  return std::max<int64_t>(16, 4 * std::max<int64_t>(tiles * N * (int64_t)C, columns * N));
}

static int launch_kl_snc(const DskdDsgfdKlArgs* a, cudaStream_t st) {
  DSKD_REQUIRE(a->d_cell_weight != nullptr && a->d_grad_rows == nullptr,
               "dsgfd_kl: the [S,N,C] layout takes per-cell masks (sg_out / fg_only), forward only");
  DSKD_REQUIRE(a->C % 4 == 0 && a->C <= 4096 && aligned16(a->d_student[0]) && aligned16(a->d_teacher[0]),
               "dsgfd_kl: [S,N,C] needs C %% 4 == 0, C <= 4096 and 16-byte aligned memory");
  KlSncParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.num_levels = a->num_levels;
  prm.N = a->N;
  prm.C = a->C;
  prm.groups = (a->C + 127) / 128;
  prm.temperature = a->temperature;
  prm.kscale = (float)(1.4426950408889634 / (double)a->temperature);
  prm.cells_per_image = a->cells_per_image;
  prm.student = a->d_student[0];
  prm.teacher = a->d_teacher[0];
  prm.cell_weight = a->d_cell_weight;
  prm.loss = a->d_loss;
  prm.redo_mask = static_cast<unsigned*>(a->d_workspace);
  int64_t tasks = 0, columns = 0;
  for (int l = 0; l < a->num_levels; ++l) {
    prm.levels[l] = a->levels[l];
    prm.scale[l] = a->scale[l];
    prm.task_start[l] = (int)tasks;
    tasks += (int64_t)a->levels[l].W * a->N * prm.groups;
    columns += (int64_t)a->levels[l].W * a->N;
  }
  DSKD_REQUIRE(tasks < (1ll << 31), "dsgfd_kl: too many columns");
  prm.task_start[a->num_levels] = (int)tasks;
  DSKD_CUDA_OK(cudaMemsetAsync(prm.redo_mask, 0, (size_t)columns * 4, st));
  const int blocks = (int)ceil_div(tasks, kSncWarps);
  dsgfd_kl_snc_kernel<<<blocks, 32 * kSncWarps, 0, st>>>(prm);
  DSKD_LAUNCH_OK("dsgfd_kl_snc_kernel");
  dsgfd_kl_snc_redo_kernel<<<kKlRedoCtas, 32 * kSncWarps, 0, st>>>(prm);
  DSKD_LAUNCH_OK("dsgfd_kl_snc_redo_kernel");
  return DSKD_OK;
}

extern "C" int dskd_dsgfd_kl_fwd_bwd(const DskdDsgfdKlArgs* a, void* stream) {
  DSKD_REQUIRE(a != nullptr, "dskd_dsgfd_kl_fwd_bwd: null args");
  DSKD_REQUIRE(a->num_levels > 0 && a->num_levels <= DSKD_MAX_LEVELS && a->N >= 0 && a->C > 0, "dsgfd_kl: bad sizes");
  DSKD_REQUIRE(a->temperature >= 1.f, "dsgfd_kl: T must be >= 1 (kd_loss.py:58)");
  const bool cell = a->d_cell_weight != nullptr;
  DSKD_REQUIRE(cell != (a->d_owner != nullptr), "dsgfd_kl: exactly one of d_owner / d_cell_weight must be set");
  DSKD_REQUIRE(a->d_loss != nullptr, "dsgfd_kl: d_loss is null");
  DSKD_REQUIRE(cell || a->num_pairs == 0 || a->d_rows != nullptr, "dsgfd_kl: d_rows is null");
  // (owner * C) * 32 + lane is one 32-bit key of the run records
  DSKD_REQUIRE(cell || (int64_t)a->num_pairs * a->C < (1ll << 26), "dsgfd_kl: num_pairs * C must stay below 2^26");
  DSKD_REQUIRE(a->layout == DSKD_LAYOUT_NCHW || a->layout == DSKD_LAYOUT_SNC, "dsgfd_kl: bad layout %d", a->layout);
  if (a->N == 0) return DSKD_OK;
  const int64_t ws_bytes = dskd_dsgfd_kl_workspace_bytes(a->N, a->num_levels, a->levels, a->C);
  DSKD_REQUIRE(a->d_workspace != nullptr && a->workspace_bytes >= ws_bytes && (reinterpret_cast<uintptr_t>(a->d_workspace) & 15u) == 0,
               "dsgfd_kl: workspace must be %lld bytes (dskd_dsgfd_kl_workspace_bytes), 16-byte aligned", (long long)ws_bytes);
  KlParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.num_levels = a->num_levels;
  prm.N = a->N;
  prm.C = a->C;
  prm.temperature = a->temperature;
  prm.kscale = (float)(1.4426950408889634 / (double)a->temperature);
  prm.cells_per_image = a->cells_per_image;
  prm.owner = a->d_owner;
  prm.rows = a->d_rows;
  prm.grad_rows = a->d_grad_rows;
  prm.cell_weight = a->d_cell_weight;
  prm.loss = a->d_loss;
  prm.redo_mask = static_cast<unsigned*>(a->d_workspace);
  int64_t cells = 0;
  int max_h = 0;
  for (int l = 0; l < a->num_levels; ++l) {
    DSKD_REQUIRE(a->levels[l].H > 0 && a->levels[l].W > 0 && a->levels[l].cell_offset == cells,
                 "dsgfd_kl: level %d is not densely packed", l);
    DSKD_REQUIRE(a->layout == DSKD_LAYOUT_SNC ? (l > 0 || (a->d_student[0] && a->d_teacher[0])) : (a->d_student[l] && a->d_teacher[l]),
                 "dsgfd_kl: null feature pointer at level %d", l);
    DSKD_REQUIRE((int64_t)a->levels[l].H * a->levels[l].W < (1ll << 31), "dsgfd_kl: level %d too large", l);
    cells += (int64_t)a->levels[l].H * a->levels[l].W;
    max_h = std::max(max_h, a->levels[l].H);
  }
  DSKD_REQUIRE(cells == a->cells_per_image, "dsgfd_kl: cells_per_image mismatch");
  cudaStream_t st = as_stream(stream);
  if (a->layout == DSKD_LAYOUT_SNC) return launch_kl_snc(a, st);
  DSKD_REQUIRE(max_h <= kKlMaxH, "dsgfd_kl: H (%d) above the supported %d", max_h, kKlMaxH);
  // tuning hook (tools/kl_perf.py): DSKD_KL_TUNE="channels_per_pass,rows_per_block,channels_per_cta,pool_records,dbg,warps,row_parts"
  const bool pair_ok = a->C % 2 == 0 && (cell || ((reinterpret_cast<uintptr_t>(a->d_rows) & 7u) == 0 &&
                                                  (reinterpret_cast<uintptr_t>(a->d_grad_rows) & 7u) == 0));
  const bool grad = !cell && a->d_grad_rows != nullptr;
  int nc = pair_ok ? 2 : 1, blk = grad ? kKlBlk : kKlBlkFwd;
  // Small batches of the gradient kernel: a tile's CTA lasts as long as its column is tall and there are too few CTAs to
  // hide that, so two warps share a column (row parts) in 16-warp CTAs (measured at 4 images, tools/kl_perf.py --tune).
  int chunk = kKlChunk, cap = kKlPool, dbg = 0, nwarps = 0;
  int rsplit = grad && a->N <= kKlSmallBatch ? 2 : 1;
  if (const char* tune = getenv("DSKD_KL_TUNE"))
    sscanf(tune, "%d,%d,%d,%d,%d,%d,%d", &nc, &blk, &chunk, &cap, &dbg, &nwarps, &rsplit);
  // one warp per `nc` channels of the chunk and row part
  if (nwarps == 0) nwarps = std::max(rsplit, std::min(kKlMaxWarps, std::min(a->C, chunk) / nc * rsplit));
  DSKD_REQUIRE((nc == 1 || (nc == 2 && pair_ok)) && chunk >= 1 && chunk <= 32 && chunk % nc == 0 && cap >= 1 &&
                   (blk == kKlBlk || blk == kKlBlkFwd) && rsplit >= 1 && rsplit <= kKlMaxSplit && nwarps >= rsplit &&
                   nwarps <= kKlMaxWarps && nwarps % rsplit == 0 && ((nwarps / rsplit) & (nwarps / rsplit - 1)) == 0,
               "dsgfd_kl: bad DSKD_KL_TUNE");
  prm.max_blocks = (max_h + blk - 1) / blk;
  prm.max_h = max_h;
  prm.chunk = std::min(a->C, chunk);
  prm.pool_cap = cap;
  prm.rsplit = rsplit;
  prm.dbg = dbg;
  const int nchunks = (a->C + prm.chunk - 1) / prm.chunk;
  int blocks = 0;
  for (int l = 0; l < a->num_levels; ++l) {
    prm.levels[l] = a->levels[l];
    prm.student[l] = a->d_student[l];
    prm.teacher[l] = a->d_teacher[l];
    prm.scale[l] = a->scale[l];
    prm.wtiles[l] = (a->levels[l].W + 31) / 32;
    prm.wpt[l] = (a->levels[l].W + prm.wtiles[l] - 1) / prm.wtiles[l];
    prm.block_start[l] = blocks;
    blocks += prm.wtiles[l] * nchunks * a->N;
  }
  prm.block_start[a->num_levels] = blocks;
  prm.num_blocks = blocks;
  size_t smem = KlSmem(prm.max_blocks, blk, grad, cap, 1 + 2 * nc, nwarps, rsplit > 1 ? 3 * nc : 0).total;
  if (smem > 227 * 1024 && nwarps > kKlWarps) {  // very tall levels: the 8-warp layout needs half the record pools
    nwarps = kKlWarps;
    prm.rsplit = rsplit = 1;
    smem = KlSmem(prm.max_blocks, blk, grad, cap, 1 + 2 * nc, nwarps, 0).total;
  }
  DSKD_REQUIRE(smem <= 227 * 1024, "dsgfd_kl: H (%d) needs %zu bytes of shared memory", max_h, smem);
  bool launched = false;
#define DSKD_KL_VARIANT(CELLV, GRADV, NCV, RV)                                                                  \
  if (!launched && nc == NCV && blk == RV) {                                                                    \
    DSKD_CUDA_OK(cudaFuncSetAttribute(dsgfd_kl_stream_kernel<CELLV, GRADV, NCV, RV>,                            \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
    dsgfd_kl_stream_kernel<CELLV, GRADV, NCV, RV><<<blocks, 32 * nwarps, smem, st>>>(prm);                      \
    launched = true;                                                                                            \
  }
#define DSKD_KL_VARIANTS(CELLV, GRADV)               \
  DSKD_KL_VARIANT(CELLV, GRADV, 2, kKlBlk)           \
  DSKD_KL_VARIANT(CELLV, GRADV, 2, kKlBlkFwd)        \
  DSKD_KL_VARIANT(CELLV, GRADV, 1, kKlBlk)           \
  DSKD_KL_VARIANT(CELLV, GRADV, 1, kKlBlkFwd)
  if (cell) { DSKD_KL_VARIANTS(true, false) }
  else if (grad) { DSKD_KL_VARIANTS(false, true) }
  else { DSKD_KL_VARIANTS(false, false) }
#undef DSKD_KL_VARIANTS
#undef DSKD_KL_VARIANT
  DSKD_REQUIRE(launched, "dsgfd_kl: no kernel variant for DSKD_KL_TUNE=%d,%d", nc, blk);
  DSKD_LAUNCH_OK("dsgfd_kl_stream_kernel");
  // the redo launch: a few loads per CTA when no block left anything behind
  const size_t redo_smem = (size_t)max_h * 32 * 4;
#define DSKD_KL_REDO(CELLV, GRADV)                                                                                        \
  do {                                                                                                                    \
    DSKD_CUDA_OK(cudaFuncSetAttribute(dsgfd_kl_redo_kernel<CELLV, GRADV>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                      (int)redo_smem));                                                                   \
    dsgfd_kl_redo_kernel<CELLV, GRADV><<<kKlRedoCtas, 32 * kKlWarps, redo_smem, st>>>(prm);                               \
  } while (0)
  if (cell) DSKD_KL_REDO(true, false);
  else if (grad) DSKD_KL_REDO(false, true);
  else DSKD_KL_REDO(false, false);
#undef DSKD_KL_REDO
  DSKD_LAUNCH_OK("dsgfd_kl_redo_kernel");
  return DSKD_OK;
}
