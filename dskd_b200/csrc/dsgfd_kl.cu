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

constexpr int kKlWarps = 8;       // warps per CTA (one channel at a time each)
// Measured on B200 (tools/kl_perf.py --tune, 16 images of 800x1333): two channels per pass, 4-row blocks loaded and consumed
// in place (no register ring), 4 CTAs x 8 warps per SM at 64 registers: 139 us; with a 2-stage ring at 127 registers
// (2 CTAs per SM) 168 us; one channel per pass (5-row blocks, 2 stages, 4 CTAs per SM) 169 us.
constexpr int kKlBlk = 4;         // rows per block: the unit of loading and of skipping rows without boxes (gradient kernels)
constexpr int kKlBlkFwd = 5;      //   ... forward-only kernels
constexpr int kKlStages = 1;      // row blocks in the register ring: the loads of kKlStages - 1 blocks are in flight ahead
constexpr int kKlMinCtas = 4;     // CTAs per SM the register budget is cut for
constexpr int kKlChunk = 16;      // channels per CTA (8 for small batches: twice the CTAs to fill 148 SMs x 4)
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
  size_t endm, anym, pool, act, total;
  // rec_words: 32-bit words per run record: key + (A, B) per channel of the pass
  __host__ __device__ KlSmem(int max_blocks, int blk, bool grad, int pool_cap, int rec_words) {
    const size_t rows = (size_t)max_blocks * blk;
    endm = rows * 32 * 4;
    anym = endm + rows * 4;
    pool = (anym + rows * 4 + 15) / 16 * 16;
    act = pool + (grad ? (size_t)kKlWarps * pool_cap * rec_words * 4 : 0);
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

// Bottom of a column for one channel of a pass: the runs parked in the pool (planes ia / ib hold A / B of this channel)
// become d loss / d mask rows: neighbouring records of one owner are added up first (segmented scan over the lanes), then
// one red.global per segment.
__device__ __forceinline__ void kl_flush_pool(const int* pool, int pool_words, int ia, int ib, int count, int lane,
                                              const float* norm /* [32][stride]: rt, rs of the source lane */, int stride,
                                              int io, float gcoef, float* __restrict__ growc, bool no_red) {
  constexpr unsigned kFull = 0xffffffffu;
  for (int j0 = 0; j0 < count; j0 += 32) {
    const int j = j0 + lane;
    const bool have = j < count;
    const int key = have ? pool[j] : -1 - lane;  // owner * C * 32 + lane of the column
    float g = 0.f;
    if (have) {
      const float* nr = norm + (key & 31) * stride + io;
      g = gcoef * (__int_as_float(pool[j + ia * pool_words]) * nr[0] - __int_as_float(pool[j + ib * pool_words]) * nr[1]);
    }
    const int own = key >> 5;
    const int prev = __shfl_up_sync(kFull, own, 1), next = __shfl_down_sync(kFull, own, 1);
    bool open = lane > 0 && prev == own;  // the segment continues to the left
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const float gv = __shfl_up_sync(kFull, g, d);
      const bool ov = __shfl_up_sync(kFull, (int)open, d);
      if (open && lane >= d) { g += gv; open = ov; }
    }
    const bool tail = lane == 31 || next != own;
    if (have && tail && !no_red) atomicAdd(growc + own, g);
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
template <bool CELL, bool GRAD, int NC, int R, int NB, int MINB>
__global__ void __launch_bounds__(32 * kKlWarps, MINB) dsgfd_kl_stream_kernel(const __grid_constant__ KlParams prm) {
  extern __shared__ __align__(16) unsigned char kl_smem[];
  __shared__ double red[32];
  __shared__ float norm_s[kKlWarps][32][2 * NC];  // per lane: (1 / sum e^y, 1 / sum e^x) of each channel of the pass
  __shared__ int nact_s, nadd_s, nends_s;
  __shared__ unsigned redo_s;
  constexpr unsigned kFull = 0xffffffffu;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const KlTile tl = kl_decode_tile(prm, blockIdx.x);
  const int lvl = tl.lvl;
  const int H = prm.levels[lvl].H, C = prm.C;
  const unsigned W = (unsigned)prm.levels[lvl].W;
  const int nblk = (H + R - 1) / R, rows_pad = nblk * R;

  constexpr int kRec = 1 + 2 * NC;  // words per run record
  const KlSmem lay(prm.max_blocks, R, GRAD, prm.pool_cap, kRec);
  KlTables tb;
  tb.moff = reinterpret_cast<int*>(kl_smem);
  tb.mw = reinterpret_cast<float*>(kl_smem);
  tb.endm = reinterpret_cast<unsigned*>(kl_smem + lay.endm);
  tb.anym = reinterpret_cast<unsigned*>(kl_smem + lay.anym);
  tb.act = reinterpret_cast<unsigned short*>(kl_smem + lay.act);

  // ---- once per CTA: owner table, run ends, list of active row blocks
  {
    const int64_t cell0 = (int64_t)tl.img * prm.cells_per_image + prm.levels[lvl].cell_offset + tl.w0;
    for (int i = tid; i < rows_pad * 32; i += 32 * kKlWarps) {
      const int h = i >> 5, l = i & 31;
      const bool in = h < H && l < tl.wn;
      if (CELL) tb.mw[i] = in ? __ldg(prm.cell_weight + cell0 + (int64_t)h * W + l) : 0.f;
      else {
        const int o = in ? __ldg(prm.owner + cell0 + (int64_t)h * W + l) : -1;
        tb.moff[i] = o >= 0 ? o * C : -1;
      }
    }
    if (tid == 0) { nends_s = 0; redo_s = 0u; }
    __syncthreads();
    // per row: lanes whose run of one owner ends here, lanes with a box at all
    int ends_here = 0;
    for (int h = warp; h < rows_pad; h += kKlWarps) {
      bool on, end = false;
      if (CELL) on = tb.mw[h * 32 + lane] != 0.f;
      else {
        const int o = tb.moff[h * 32 + lane];
        const int nx = (h + 1 < rows_pad) ? tb.moff[(h + 1) * 32 + lane] : -1;
        on = o >= 0;
        end = on && nx != o;
      }
      const unsigned am = __ballot_sync(kFull, on), em = __ballot_sync(kFull, end);
      if (lane == 0) { tb.anym[h] = am; tb.endm[h] = em; }
      ends_here += __popc(em);
    }
    if (GRAD && lane == 0 && ends_here) atomicAdd(&nends_s, ends_here);
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
    }
    __syncthreads();
  }
  const int nact = nact_s;
  if (nact == 0) {  // no box touches this tile: every column's KL is exactly 0
    if (tid == 0) prm.redo_mask[blockIdx.x] = 0u;
    return;
  }
  const int c_begin = tl.chunk * prm.chunk, c_end = min(C, c_begin + prm.chunk);
  // warps working side by side: each needs pool room for every run of the tile
  int share = 1;
  if (GRAD) {
    const int nends = nends_s;
    while (share < kKlWarps && nends > share * prm.pool_cap) share *= 2;
    if (nends > share * prm.pool_cap) {  // not even one warp with every pool: the whole block is left to the redo launch
      if (tid == 0) prm.redo_mask[blockIdx.x] = 0xffffffffu >> (32 - (c_end - c_begin));
      return;
    }
  }
  const int wstep = kKlWarps / share;  // active warps: 0 .. wstep - 1, warp w owns the pools w * share ..
  // warp w's records: planes (key, A, B per channel) of share * pool_cap words each
  const int pool_words = share * prm.pool_cap;
  const int* pool = reinterpret_cast<const int*>(kl_smem + lay.pool) + (size_t)warp * kRec * pool_words;
  const unsigned pool_addr = (unsigned)__cvta_generic_to_shared(pool);
  double kl_total = 0.0;

  if (warp < wstep) {
    const bool col_ok = lane < tl.wn;
    for (int c = c_begin + NC * warp; c < c_end; c += NC * wstep) {
      const int64_t plane = ((int64_t)tl.img * C + c) * ((int64_t)H * W) + tl.w0 + lane;
      bool ok;
      int base;
      float kl = 0.f;
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
        kl_stream_pass2<CELL, GRAD, R, NB>(tb, ch, W, lane, nact, prm.kscale, pool_addr, pool_words, prm.dbg);
        const float nadd = (float)nadd_s;
        const float ss0 = ch.ss.x + nadd, st0 = ch.st.x + nadd, ss1 = ch.ss.y + nadd, st1 = ch.st.y + nadd;
        ok = __all_sync(kFull, !col_ok || (kl_in_range(ss0, st0, ch.ws.x) && kl_in_range(ss1, st1, ch.ws.y)));
        base = ch.base;
        if (ok) {
          kl = kl_col(ss0, st0, ch.ws.x) + kl_col(ss1, st1, ch.ws.y);
          if (GRAD) {
            norm_s[warp][lane][0] = __fdividef(1.f, st0);
            norm_s[warp][lane][1] = __fdividef(1.f, ss0);
            norm_s[warp][lane][2] = __fdividef(1.f, st1);
            norm_s[warp][lane][3] = __fdividef(1.f, ss1);
          }
        }
      } else {
        KlChan ch;
        ch.Sp = prm.student[lvl] + plane;
        ch.Tp = prm.teacher[lvl] + plane;
        ch.rowsc = CELL ? nullptr : prm.rows + c;
        asm volatile("" : "+l"(ch.Sp), "+l"(ch.Tp), "+l"(ch.rowsc));
        ch.ss = ch.st = ch.ws = 0.f;
        ch.base = 0;
        kl_stream_pass<CELL, GRAD, R, NB>(tb, ch, W, lane, nact, prm.kscale, pool_addr, pool_words, prm.dbg);
        const float nadd = (float)nadd_s;
        const float ss = ch.ss + nadd, st = ch.st + nadd;
        ok = __all_sync(kFull, !col_ok || kl_in_range(ss, st, ch.ws));
        base = ch.base;
        if (ok) {
          kl = kl_col(ss, st, ch.ws);
          if (GRAD) {
            norm_s[warp][lane][0] = __fdividef(1.f, st);
            norm_s[warp][lane][1] = __fdividef(1.f, ss);
          }
        }
      }
      // ---- bottom of the column: KL, parked runs
      if (ok) {
        if (col_ok) kl_total += (double)kl;
        if (GRAD && !(prm.dbg & 4)) {
          const float gcoef = prm.scale[lvl] * prm.temperature / (float)H;  // d loss / d pred = scale * (T/H) * (p - q)
          __syncwarp();
#pragma unroll
          for (int k = 0; k < NC; ++k)
            kl_flush_pool(pool, pool_words, 1 + 2 * k, 2 + 2 * k, base, lane, &norm_s[warp][0][0], 2 * NC, 2 * k, gcoef,
                          prm.grad_rows + c + k, prm.dbg & 1);
          __syncwarp();
        }
      } else if (lane == 0) {
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
  // tuning hook (tools/kl_perf.py): DSKD_KL_TUNE="channels_per_pass,rows_per_block,stages,ctas_per_sm,channels_per_cta,pool_records,dbg"
  const bool pair_ok = a->C % 2 == 0 && (cell || (reinterpret_cast<uintptr_t>(a->d_rows) & 7u) == 0);
  const bool grad = !cell && a->d_grad_rows != nullptr;
  int nc = pair_ok ? 2 : 1, blk = grad ? kKlBlk : kKlBlkFwd, stages = kKlStages, minb = kKlMinCtas;
  int chunk = a->N <= 4 ? kKlChunk / 2 : kKlChunk, cap = kKlPool, dbg = 0;
  if (const char* tune = getenv("DSKD_KL_TUNE"))
    sscanf(tune, "%d,%d,%d,%d,%d,%d,%d", &nc, &blk, &stages, &minb, &chunk, &cap, &dbg);
  DSKD_REQUIRE((nc == 1 || (nc == 2 && pair_ok)) && chunk >= 1 && chunk <= 32 && chunk % nc == 0 && cap >= 1,
               "dsgfd_kl: bad DSKD_KL_TUNE");
  prm.max_blocks = (max_h + blk - 1) / blk;
  prm.max_h = max_h;
  prm.chunk = std::min(a->C, chunk);
  prm.pool_cap = cap;
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
  const size_t smem = KlSmem(prm.max_blocks, blk, grad, cap, 1 + 2 * nc).total;
  DSKD_REQUIRE(smem <= 227 * 1024, "dsgfd_kl: H (%d) needs %zu bytes of shared memory", max_h, smem);
  bool launched = false;
#define DSKD_KL_VARIANT(CELLV, GRADV, NCV, RV, NBV, MB)                                                         \
  if (!launched && nc == NCV && blk == RV && stages == NBV && minb == MB) {                                     \
    DSKD_CUDA_OK(cudaFuncSetAttribute(dsgfd_kl_stream_kernel<CELLV, GRADV, NCV, RV, NBV, MB>,                   \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
    dsgfd_kl_stream_kernel<CELLV, GRADV, NCV, RV, NBV, MB><<<blocks, 32 * kKlWarps, smem, st>>>(prm);           \
    launched = true;                                                                                            \
  }
#define DSKD_KL_VARIANTS(CELLV, GRADV)          \
  DSKD_KL_VARIANT(CELLV, GRADV, 2, 4, 1, 4)     \
  DSKD_KL_VARIANT(CELLV, GRADV, 2, 5, 1, 4)     \
  DSKD_KL_VARIANT(CELLV, GRADV, 2, 5, 1, 3)     \
  DSKD_KL_VARIANT(CELLV, GRADV, 2, 4, 2, 3)     \
  DSKD_KL_VARIANT(CELLV, GRADV, 2, 5, 2, 2)     \
  DSKD_KL_VARIANT(CELLV, GRADV, 1, 4, 1, 4)     \
  DSKD_KL_VARIANT(CELLV, GRADV, 1, 5, 1, 4)     \
  DSKD_KL_VARIANT(CELLV, GRADV, 1, 5, 2, 4)
  if (cell) { DSKD_KL_VARIANTS(true, false) }
  else if (grad) { DSKD_KL_VARIANTS(false, true) }
  else { DSKD_KL_VARIANTS(false, false) }
#undef DSKD_KL_VARIANTS
#undef DSKD_KL_VARIANT
  DSKD_REQUIRE(launched, "dsgfd_kl: no kernel variant for DSKD_KL_TUNE=%d,%d,%d,%d", nc, blk, stages, minb);
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
