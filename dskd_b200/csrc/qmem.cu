// Query x memory contraction on tcgen05 tensor cores fed by TMA (north_star's
// "query x memory mask"): per image  score[s, j] = <memory[s, i, :], hs_T[keepid_j, :]>  for the S ~ 22k
// encoder tokens and the K_i matched teacher queries, never written to HBM -- the epilogue turns each
// token's row of scores into ONE cell weight straight out of tensor memory:
//
//   z_sj = score[s, j] / (sqrt(C) * temp)
//   w_s  = sqrt( sum_j c_j e^{z_sj} / (1 + sum_j e^{z_sj}) )        c_j = teacher confidence of detection j
//
// i.e. a softmax over the matched queries with a null "background" logit 0 (soft ownership instead of the
// reference's hard last-writer-wins rectangles, head_il.py:688-706), square-rooted like the reference's other
// cell masks because the MSE squares it (head_il.py:914,1119).  The reference has no such contraction
// (SURVEY.md section 0.4); this is the unpinned extension row, its oracle is oracle/qmem.py.
//
// Kernel anatomy (one CTA per SM, persistent, 320 threads, cta_group::1):
//   warp 0   TMA producer : memory tiles [128 tokens x 32 ch] fp32 (one 128-byte swizzle row per token) through a
//            ring of kAStages shared-memory stages; the query block [NB x C] of the current image is loaded
//            once and stays resident in shared memory (it is the B operand of every tile of that image)
//   warp 1   MMA issuer   : tcgen05.mma kind::tf32, M = 128 tokens, N = NB queries, K = 8 per instruction,
//            fp32 accumulators in TMEM, kAccStages accumulator stages of kMaxNB columns
//   warps 2-9 epilogue    : two sets of 4 warps alternating over the tiles; tcgen05.ld 32 lanes x 32 columns,
//            online softmax in registers (thread = token), one coalesced 4-byte store per token
// More than kMaxNB matched queries per image are split into query blocks handled by neighbouring CTAs (the
// memory tile is then read from HBM once and from L2 nblk times) and merged by qmem_combine_kernel.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace dskd {

constexpr int kTokTile = 128;             // UMMA M: tokens per tile = TMEM lanes
constexpr int kSlabCh = 32;               // fp32 channels per 128-byte swizzle row = one TMA box row
constexpr int kMaxNB = 160;               // queries per block: UMMA N (multiple of 16) and TMEM columns per stage
constexpr int kAStages = 4;               // memory-tile ring
constexpr int kAccStages = 3;             // 3 x 160 = 480 of the 512 TMEM columns
constexpr int kTmemCols = 512;
constexpr int kEpiSets = 2;               // epilogue warp sets (4 warps each) alternating over the tiles
constexpr int kQmemThreads = 64 + 128 * kEpiSets;
constexpr uint32_t kAStageBytes = kTokTile * kSlabCh * 4;  // 16 KB

struct QmemParams {
  int N, C, S, nblk, NB, tiles_per_image;
  long long total_tiles;
  float score_scale;        // log2(e) / (sqrt(C) * temp)
  const int* box_start;     // [N+1]
  const float* cpad;        // [N, nblk, NB] confidences, zero padded
  float* part;              // [N*nblk][3][S] (max, den, num) when nblk > 1
  float* weight;            // [N, S]
  int debug;                // DSKD_QMEM_DEBUG bits (perf experiments): 1 skip epilogue math, 2 skip MMA issue
};

// ------------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]^T, both operands K-major tf32 (fp32 bit patterns, low mantissa bits ignored)
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when they complete (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major, SWIZZLE_128B: rows of 128 bytes, 8-row swizzle atoms 1024 bytes apart
// (SBO), LBO unused, descriptor version 1 (sm_100), base offset 0 (tiles are 1024-byte aligned).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// ------------------------------------------------------------------------------------------------ kernel
// dynamic shared memory (1024-byte aligned): [Q slabs: C/32 x NB x 128 B][A ring: kAStages x 16 KB][barriers]
__global__ void __launch_bounds__(kQmemThreads, 1)
qmem_weight_kernel(const __grid_constant__ CUtensorMap tmap_mem, const __grid_constant__ CUtensorMap tmap_q,
                   const __grid_constant__ QmemParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int num_slabs = p.C / kSlabCh;
  const uint32_t q_slab_bytes = (uint32_t)p.NB * 128u;
  const uint32_t smem_q = smem_base;
  const uint32_t smem_a = smem_q + (uint32_t)num_slabs * q_slab_bytes;
  const uint32_t bars = smem_a + kAStages * kAStageBytes;
  // barrier slots (8 bytes each)
  const uint32_t bar_full = bars, bar_empty = bars + 8 * kAStages, bar_qfull = bars + 16 * kAStages,
                 bar_qempty = bar_qfull + 8, bar_accfull = bar_qfull + 16, bar_accempty = bar_accfull + 8 * kAccStages,
                 tmem_slot = bar_accempty + 8 * kAccStages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // work of this CTA: query block b of a contiguous range of (image, token tile) pairs
  const int b = blockIdx.x % p.nblk;
  const long long g = blockIdx.x / p.nblk, G = gridDim.x / p.nblk;
  const long long t_begin = g * p.total_tiles / G, t_end = (g + 1) * p.total_tiles / G;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kAStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_qfull, 1);
    mbar_init(bar_qempty, 1);
    for (int s = 0; s < kAccStages; ++s) { mbar_init(bar_accfull + 8 * s, 1); mbar_init(bar_accempty + 8 * s, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: whole warp allocates, address lands in shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, qe_phase = 0;
      int cur_img = -1;
      for (long long t = t_begin; t < t_end; ++t) {
        const int img = (int)(t / p.tiles_per_image), tt = (int)(t % p.tiles_per_image);
        if (img != cur_img) {
          if (cur_img >= 0) { mbar_wait(bar_qempty, qe_phase); qe_phase ^= 1; }  // MMAs on the old block are done
          mbar_arrive_expect_tx(bar_qfull, (uint32_t)num_slabs * q_slab_bytes);
          for (int ks = 0; ks < num_slabs; ++ks)
            tma_load_2d(&tmap_q, bar_qfull, smem_q + ks * q_slab_bytes, ks * kSlabCh, (img * p.nblk + b) * p.NB);
          cur_img = img;
        }
        for (int ks = 0; ks < num_slabs; ++ks) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          mbar_arrive_expect_tx(bar_full + 8 * stage, kAStageBytes);
          tma_load_3d(&tmap_mem, bar_full + 8 * stage, smem_a + stage * kAStageBytes, ks * kSlabCh, img, tt * kTokTile);
          if (++stage == kAStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      // instruction descriptor: D fp32, A/B tf32, both K-major, N = NB, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.NB >> 3) << 17) | ((uint32_t)(kTokTile >> 4) << 24);
      uint32_t stage = 0, phase = 0, qf_phase = 0, acc = 0, acc_phase = 0;
      int cur_img = -1;
      for (long long t = t_begin; t < t_end; ++t) {
        const int img = (int)(t / p.tiles_per_image);
        if (img != cur_img) { mbar_wait(bar_qfull, qf_phase); qf_phase ^= 1; cur_img = img; }
        mbar_wait(bar_accempty + 8 * acc, acc_phase ^ 1);  // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kMaxNB;
        for (int ks = 0; ks < num_slabs; ++ks) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint64_t da = smem_desc_sw128(smem_a + stage * kAStageBytes);
          const uint64_t db = smem_desc_sw128(smem_q + ks * q_slab_bytes);
#pragma unroll
          for (int kk = 0; kk < kSlabCh / 8; ++kk)  // 8 tf32 = 32 bytes along K per instruction: +2 in 16-byte units
            if (!(p.debug & 2)) umma_tf32(tmem_d, da + 2 * kk, db + 2 * kk, idesc, (ks | kk) ? 1u : 0u);
          umma_commit(bar_empty + 8 * stage);  // frees the memory-tile stage when those MMAs retire
          if (++stage == kAStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(bar_accfull + 8 * acc);  // accumulator ready for the epilogue
        const bool last_of_image = (t + 1 == t_end) || ((int)((t + 1) / p.tiles_per_image) != img);
        if (last_of_image) umma_commit(bar_qempty);  // the resident query block may be replaced
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ================================================================ epilogue: thread = token (TMEM lane)
    // Two sets of four warps alternate over the tiles, so each SM sub-partition always has two epilogue warps to
    // interleave (a single warp per sub-partition is bound by its own dependent-issue latency).
    const int sub = warp & 3;               // TMEM sub-partition this warp may read: lanes [32*sub, 32*sub + 32)
    const int eset = (warp - 2) >> 2;
    const bool single = p.nblk == 1;
    const float scale = p.score_scale;
    long long n = eset;                     // ordinal of the tile inside this CTA: fixes the accumulator stage
    for (long long t = t_begin + eset; t < t_end; t += kEpiSets, n += kEpiSets) {
      const uint32_t acc = (uint32_t)(n % kAccStages), acc_phase = (uint32_t)((n / kAccStages) & 1);
      const int img = (int)(t / p.tiles_per_image), tt = (int)(t % p.tiles_per_image);
      const int K = p.box_start[img + 1] - p.box_start[img];
      const int kv = (p.debug & 1) ? 0 : max(0, min(p.NB, K - b * p.NB));  // valid query columns of this block
      const float4* __restrict__ cj = reinterpret_cast<const float4*>(p.cpad + ((long long)img * p.nblk + b) * p.NB);
      mbar_wait(bar_accfull + 8 * acc, acc_phase);
      tc_fence_after();
      // the null logit 0 is part of the softmax; with several query blocks it is added once, by the combine kernel
      float mx = single ? 0.f : -1e30f, den0 = single ? 1.f : 0.f, den1 = 0.f, num0 = 0.f, num1 = 0.f;
      const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + acc * kMaxNB;
      for (int c0 = 0; c0 < kv; c0 += 32) {
        float c[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {  // confidences of these 32 columns: warp-uniform 128-bit loads (L1 broadcast)
          const float4 c4 = __ldg(cj + (c0 >> 2) + q);
          c[4 * q] = c4.x; c[4 * q + 1] = c4.y; c[4 * q + 2] = c4.z; c[4 * q + 3] = c4.w;
        }
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        tmem_ld_wait();
        const int nc = kv - c0;
        if (nc < 32) {  // padded query rows are zero vectors (score 0): take them out of the softmax
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j >= nc) v[j] = 0xff800000u;
        }
        float vm = __uint_as_float(v[0]);
#pragma unroll
        for (int j = 1; j < 32; ++j) vm = fmaxf(vm, __uint_as_float(v[j]));
        const float cm = vm * scale;
        if (cm > mx) {
          const float r = ex2_approx(mx - cm);
          den0 *= r; den1 *= r; num0 *= r; num1 *= r;
          mx = cm;
        }
        const float nmx = -mx;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float e0 = ex2_approx(fmaf(__uint_as_float(v[j]), scale, nmx));
          const float e1 = ex2_approx(fmaf(__uint_as_float(v[j + 1]), scale, nmx));
          den0 += e0;
          den1 += e1;
          num0 = fmaf(c[j], e0, num0);
          num1 = fmaf(c[j + 1], e1, num1);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_accempty + 8 * acc);
      const float den = den0 + den1, num = num0 + num1;
      const int tok = tt * kTokTile + sub * 32 + lane;
      if (tok < p.S) {
        if (single) {
          p.weight[(long long)img * p.S + tok] = sqrtf(__fdividef(num, den));
        } else {
          float* o = p.part + ((long long)img * p.nblk + b) * 3 * p.S + tok;
          o[0] = mx;
          o[p.S] = den;
          o[2 * (long long)p.S] = num;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// Query rows of every (image, block), zero padded to NB rows: Qpad[N, nblk, NB, C] and cpad[N, nblk, NB].
__global__ void __launch_bounds__(256) qmem_gather_kernel(const float* __restrict__ hs_teacher, const int64_t* __restrict__ keepid,
                                                          const float* __restrict__ scores, const int* __restrict__ box_start,
                                                          int nblk, int NB, int C, int64_t num_rows, float* __restrict__ qpad,
                                                          float* __restrict__ cpad) {
  const int r = blockIdx.x;  // (img * nblk + b) * NB + j
  const int j = r % NB, ib = r / NB, b = ib % nblk, img = ib / nblk;
  const int q = b * NB + j, K = box_start[img + 1] - box_start[img];
  const bool valid = q < K;
  int64_t src = 0;
  if (valid) {
    src = keepid[box_start[img] + q];
    if (src < 0 || src >= num_rows) src = 0;  // defensive: never read outside hs_teacher
  }
  const float4* s4 = reinterpret_cast<const float4*>(hs_teacher + src * C);
  float4* d4 = reinterpret_cast<float4*>(qpad + (int64_t)r * C);
  for (int c = threadIdx.x; c < C / 4; c += blockDim.x) d4[c] = valid ? __ldg(s4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  if (threadIdx.x == 0) cpad[r] = valid ? (scores ? scores[box_start[img] + q] : 1.f) : 0.f;
}

// Merge the per-block (max, den, num) partials and add the null logit once.
__global__ void __launch_bounds__(256) qmem_combine_kernel(const float* __restrict__ part, int N, int nblk, int S,
                                                           float* __restrict__ weight) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)N * S) return;
  const int img = (int)(idx / S), tok = (int)(idx % S);
  float M = 0.f;
  for (int b = 0; b < nblk; ++b) M = fmaxf(M, part[((int64_t)img * nblk + b) * 3 * S + tok]);
  float den = exp2f(-M), num = 0.f;
  for (int b = 0; b < nblk; ++b) {
    const float* o = part + ((int64_t)img * nblk + b) * 3 * S + tok;
    const float r = exp2f(o[0] - M);
    den = fmaf(o[S], r, den);
    num = fmaf(o[2 * (int64_t)S], r, num);
  }
  weight[idx] = sqrtf(__fdividef(num, den));
}

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

struct QmemPlan {
  int nblk, NB;
  int64_t qpad, cpad, part, total;
  QmemPlan(int N, int64_t S, int C, int kmax) {
    nblk = std::max(1, (kmax + kMaxNB - 1) / kMaxNB);
    const int per = (std::max(kmax, 1) + nblk - 1) / nblk;
    NB = (per + 15) / 16 * 16;
    int64_t off = 0;
    auto up = [](int64_t x) { return (x + 255) / 256 * 256; };
    qpad = off; off += up((int64_t)N * nblk * NB * C * 4);
    cpad = off; off += up((int64_t)N * nblk * NB * 4);
    part = off; off += (nblk > 1) ? up((int64_t)N * nblk * 3 * S * 4) : 0;
    total = off;
  }
};
}  // namespace

}  // namespace dskd

using namespace dskd;

extern "C" int64_t dskd_qmem_workspace_bytes(int32_t N, int64_t S, int32_t C, int32_t max_per_image) {
  if (N < 0 || S < 0 || C <= 0 || max_per_image < 0) return -1;
  return QmemPlan(N, S, C, max_per_image).total;
}

extern "C" int dskd_qmem_cell_weights(const DskdQmemArgs* a, void* stream) {
  DSKD_REQUIRE(a != nullptr, "dskd_qmem_cell_weights: null args");
  DSKD_REQUIRE(a->N >= 0 && a->S > 0 && a->C > 0 && a->num_pairs >= 0 && a->max_per_image >= 0, "dskd_qmem_cell_weights: bad sizes");
  DSKD_REQUIRE(a->C % kSlabCh == 0 && a->C <= 256, "dskd_qmem_cell_weights: C (%d) must be a multiple of 32 and <= 256", a->C);
  DSKD_REQUIRE(a->temperature > 0.f, "dskd_qmem_cell_weights: temperature must be positive");
  DSKD_REQUIRE(a->d_cell_weight != nullptr, "dskd_qmem_cell_weights: d_cell_weight is null");
  cudaStream_t st = as_stream(stream);
  if (a->N == 0) return DSKD_OK;
  if (a->num_pairs == 0 || a->max_per_image == 0) {  // no matched query anywhere: every weight is sqrt(0 / 1)
    DSKD_CUDA_OK(cudaMemsetAsync(a->d_cell_weight, 0, sizeof(float) * (size_t)a->N * a->S, st));
    return DSKD_OK;
  }
  DSKD_REQUIRE(a->d_memory && a->d_hs_teacher && a->d_keepid && a->d_box_start, "dskd_qmem_cell_weights: null pointer");
  DSKD_REQUIRE(aligned16(a->d_memory) && aligned16(a->d_hs_teacher), "dskd_qmem_cell_weights: tensors must be 16-byte aligned");
  const QmemPlan plan(a->N, a->S, a->C, a->max_per_image);
  DSKD_REQUIRE(a->d_workspace != nullptr && a->workspace_bytes >= plan.total &&
                   (reinterpret_cast<uintptr_t>(a->d_workspace) % 256) == 0,
               "dskd_qmem_cell_weights: workspace must be %lld bytes, 256-byte aligned", (long long)plan.total);
  char* base = static_cast<char*>(a->d_workspace);
  float* qpad = reinterpret_cast<float*>(base + plan.qpad);
  float* cpad = reinterpret_cast<float*>(base + plan.cpad);
  float* part = plan.nblk > 1 ? reinterpret_cast<float*>(base + plan.part) : nullptr;
  const int rows = a->N * plan.nblk * plan.NB;
  qmem_gather_kernel<<<rows, 64, 0, st>>>(a->d_hs_teacher, a->d_keepid, a->d_scores, a->d_box_start, plan.nblk, plan.NB, a->C,
                                          a->num_query_rows, qpad, cpad);
  DSKD_LAUNCH_OK("qmem_gather_kernel");

  EncodeTiledFn encode = encode_tiled_fn();
  if (encode == nullptr) {
    set_error("dskd_qmem_cell_weights: cuTensorMapEncodeTiled is not available from this driver");
    return DSKD_ECUDA;
  }
  CUtensorMap tmap_mem, tmap_q;
  {
    // memory [S, N, C] fp32: dims (C, N, S) fastest first; box = 32 channels x 1 image x 128 tokens
    const cuuint64_t dims[3] = {(cuuint64_t)a->C, (cuuint64_t)a->N, (cuuint64_t)a->S};
    const cuuint64_t strides[2] = {(cuuint64_t)a->C * 4, (cuuint64_t)a->N * a->C * 4};
    const cuuint32_t box[3] = {kSlabCh, 1, kTokTile};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode(&tmap_mem, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(a->d_memory), dims, strides, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("dskd_qmem_cell_weights: cuTensorMapEncodeTiled(memory) failed with CUresult %d", (int)r);
      return DSKD_ECUDA;
    }
  }
  {
    // padded queries [N*nblk*NB, C]: box = 32 channels x NB rows
    const cuuint64_t dims[2] = {(cuuint64_t)a->C, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)a->C * 4};
    const cuuint32_t box[2] = {kSlabCh, (cuuint32_t)plan.NB};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(&tmap_q, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, qpad, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("dskd_qmem_cell_weights: cuTensorMapEncodeTiled(queries) failed with CUresult %d", (int)r);
      return DSKD_ECUDA;
    }
  }
  QmemParams p;
  p.N = a->N; p.C = a->C; p.S = (int)a->S; p.nblk = plan.nblk; p.NB = plan.NB;
  p.tiles_per_image = (int)ceil_div(a->S, kTokTile);
  p.total_tiles = (long long)p.tiles_per_image * a->N;
  p.score_scale = 1.4426950408889634f / (sqrtf((float)a->C) * a->temperature);
  p.box_start = a->d_box_start;
  p.cpad = cpad;
  p.part = part;
  p.weight = a->d_cell_weight;
  { const char* dbg = getenv("DSKD_QMEM_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }
  const size_t smem = 1024 + (size_t)plan.NB * a->C * 4 + (size_t)kAStages * kAStageBytes + 256;
  DSKD_CUDA_OK(cudaFuncSetAttribute(qmem_weight_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long groups = std::max(1ll, std::min<long long>(kNumSMs / plan.nblk, p.total_tiles));
  qmem_weight_kernel<<<(unsigned)(groups * plan.nblk), kQmemThreads, smem, st>>>(tmap_mem, tmap_q, p);
  DSKD_LAUNCH_OK("qmem_weight_kernel");
  if (plan.nblk > 1) {
    const int64_t total = (int64_t)a->N * a->S;
    qmem_combine_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(part, a->N, plan.nblk, (int)a->S, a->d_cell_weight);
    DSKD_LAUNCH_OK("qmem_combine_kernel");
  }
  return DSKD_OK;
}
