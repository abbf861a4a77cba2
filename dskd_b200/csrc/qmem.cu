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
// (SURVEY.md section 0.4); this is the unpinned extension row, its definition is the qmem module of the test oracle.
//
// Kernel anatomy (one CTA per SM, persistent, 320 threads, cta_group::1):
//   warp 0   TMA producer : memory tiles [128 tokens x 32 ch] fp32 (one 128-byte swizzle row per token) through a
//            ring of up to 12 shared-memory stages (whatever fits beside the queries); the query block [NB x C] of the current image is loaded
//            once and stays resident in shared memory (it is the B operand of every tile of that image)
//   warp 1   MMA issuer   : tcgen05.mma kind::tf32, M = 128 tokens, N = NB queries, K = 8 per instruction,
//            fp32 accumulators in TMEM, kAccStages accumulator stages of kMaxNB columns
//   warps 2-9 epilogue    : two sets of 4 warps alternating over the tiles; tcgen05.ld 32 lanes x 32 columns,
//            online softmax in registers (thread = token), one coalesced 4-byte store per token
// The producer and issuer loops are run by their whole warp with the TMA / tcgen05 instructions under elect.sync, so
// descriptors and barrier addresses live in uniform registers and the UTCHMMAs of a k-slab issue back to back (see
// elect_one()).
// More than kMaxNB matched queries per image go to the CTA-pair kernel below (cta_group::2, up to 320 queries per pass
// over the memory); beyond that the queries are split into blocks handled by neighbouring CTAs or CTA pairs (the memory
// tile is then read from HBM once and from L2 nblk times) and merged by qmem_combine_kernel.
// Measured dead ends (tools/qmem_perf.py, profiles/r1/qmem_notes.md): an L2 prefetch ahead of the ring raised DRAM
// traffic by 60 % and cost 30 %; TMA multicast of the tile to the query blocks of a cluster was 10-20 % slower than
// letting each CTA fetch it (the lock-step release of a stage by every CTA outweighs the saved L2 requests).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace dskd {

constexpr int kTokTile = 128;             // UMMA M: tokens per tile = TMEM lanes
constexpr int kSlabCh = 32;               // fp32 channels per 128-byte swizzle row = one TMA box row
constexpr int kMaxNB = 160;               // queries per block: UMMA N (multiple of 16) and TMEM columns per stage
constexpr int kMaxAStages = 12;           // memory-tile ring: as many 16 KB stages as fit beside the resident queries
constexpr int kAccStages = 3;             // 3 x 160 = 480 of the 512 TMEM columns
constexpr int kTmemCols = 512;
constexpr int kEpiSets = 2;               // epilogue warp sets (4 warps each) alternating over the tiles
constexpr int kPairSlots = 4;             // accumulator slots of the CTA-pair kernel (512 columns / slot width, at most 4)
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
  int stages;               // memory-tile ring depth (<= kMaxAStages)
  int n_lo, n_hi;           // pair mode: columns of the two MMAs per k-step (n_hi may be 0); NB = n_lo + n_hi
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

// One lane of a converged warp.  The tcgen05 / TMA issue loops are run by the WHOLE warp with the instructions under this
// predicate: descriptors and barrier addresses then stay in uniform registers.  Issued from inside an `if (lane == 0)`
// branch instead, every UTCHMMA / UTMALDG sits in an ELECT + R2UR.BROADCAST "waterfall" loop of ~16 dependent
// instructions, and that loop -- not the tensor pipe -- sets the time of a k-step (measured: ~250 cycles per step).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// ---- CTA-pair (cta_group::2) variants
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 rem;\n\t"
      "mapa.shared::cluster.u32 rem, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [rem];\n\t"
      "}" ::"r"(bar), "r"(cta)
      : "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a pair: addresses the leader's barrier
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// M = 256 over the CTA pair: each CTA supplies its 128 rows of A and half of the rows of B; issued by the leader only
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs of the pair when the leader's previously issued MMAs complete
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}

// One 32-column chunk of the online softmax of a token row (thread = token): v = raw scores from TMEM, c = confidences.
__device__ __forceinline__ void softmax_chunk(uint32_t (&v)[32], const float4* __restrict__ c4p, int nc, float scale, float& mx,
                                              float& den0, float& den1, float& num0, float& num1) {
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
  for (int q = 0; q < 8; ++q) {  // confidences of these 32 columns: warp-uniform 128-bit loads (L1 broadcast)
    const float4 c4 = __ldg(c4p + q);
    const float e0 = ex2_approx(fmaf(__uint_as_float(v[4 * q]), scale, nmx));
    const float e1 = ex2_approx(fmaf(__uint_as_float(v[4 * q + 1]), scale, nmx));
    const float e2 = ex2_approx(fmaf(__uint_as_float(v[4 * q + 2]), scale, nmx));
    const float e3 = ex2_approx(fmaf(__uint_as_float(v[4 * q + 3]), scale, nmx));
    den0 += e0;
    den1 += e1;
    num0 = fmaf(c4.x, e0, num0);
    num1 = fmaf(c4.y, e1, num1);
    den0 += e2;
    den1 += e3;
    num0 = fmaf(c4.z, e2, num0);
    num1 = fmaf(c4.w, e3, num1);
  }
}

// Shared-memory matrix descriptor, K-major, SWIZZLE_128B: rows of 128 bytes, 8-row swizzle atoms 1024 bytes apart
// (SBO), LBO unused, descriptor version 1 (sm_100), base offset 0 (tiles are 1024-byte aligned).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// ------------------------------------------------------------------------------------------------ kernel
// dynamic shared memory (1024-byte aligned): [Q slabs: C/32 x NB x 128 B][A ring: p.stages x 16 KB][barriers]
__global__ void __launch_bounds__(kQmemThreads, 1)
qmem_weight_kernel(const __grid_constant__ CUtensorMap tmap_mem, const __grid_constant__ CUtensorMap tmap_q,
                   const __grid_constant__ QmemParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int num_slabs = p.C / kSlabCh;
  const uint32_t q_slab_bytes = (uint32_t)p.NB * 128u;
  const uint32_t smem_q = smem_base;
  const uint32_t smem_a = smem_q + (uint32_t)num_slabs * q_slab_bytes;
  const int num_stages = p.stages;
  const uint32_t bars = smem_a + num_stages * kAStageBytes;
  // barrier slots (8 bytes each)
  const uint32_t bar_full = bars, bar_empty = bars + 8 * kMaxAStages, bar_qfull = bars + 16 * kMaxAStages,
                 bar_qempty = bar_qfull + 8, bar_accfull = bar_qfull + 16, bar_accempty = bar_accfull + 8 * kAccStages,
                 tmem_slot = bar_accempty + 8 * kAccStages;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // warp index, provably uniform

  // work of this CTA: query block b of a contiguous range of (image, token tile) pairs
  const int b = blockIdx.x % p.nblk;
  const long long g = blockIdx.x / p.nblk, G = gridDim.x / p.nblk;
  const int t_begin = (int)(g * p.total_tiles / G), t_end = (int)((g + 1) * p.total_tiles / G);
  // the loops below walk (image, tile in image) incrementally: no division on the issue paths
  const int img_begin = t_begin / p.tiles_per_image, tt_begin = t_begin % p.tiles_per_image;

  if (threadIdx.x == 0) {
    for (int s = 0; s < num_stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
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
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);

  if (warp == 0) {
    // ================================================================ TMA producer
    uint32_t stage = 0, phase = 0, qe_phase = 0;
    int cur_img = -1;
    int img = img_begin, tt = tt_begin;
    for (int t = t_begin; t < t_end; ++t) {
      if (img != cur_img) {
        if (cur_img >= 0) { mbar_wait(bar_qempty, qe_phase); qe_phase ^= 1; }  // MMAs on the old block are done
        if (elect_one()) {
          mbar_arrive_expect_tx(bar_qfull, (uint32_t)num_slabs * q_slab_bytes);
          for (int ks = 0; ks < num_slabs; ++ks)
            tma_load_2d(&tmap_q, bar_qfull, smem_q + ks * q_slab_bytes, ks * kSlabCh, (img * p.nblk + b) * p.NB);
        }
        __syncwarp();
        cur_img = img;
      }
      for (int ks = 0; ks < num_slabs; ++ks) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(bar_full + 8 * stage, kAStageBytes);
          tma_load_3d(&tmap_mem, bar_full + 8 * stage, smem_a + stage * kAStageBytes, ks * kSlabCh, img, tt * kTokTile);
        }
        __syncwarp();
        if (++stage == num_stages) { stage = 0; phase ^= 1; }
      }
      if (++tt == p.tiles_per_image) { tt = 0; ++img; }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (whole warp, one elected lane issues)
    // instruction descriptor: D fp32, A/B tf32, both K-major, N = NB, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.NB >> 3) << 17) | ((uint32_t)(kTokTile >> 4) << 24);
    uint32_t stage = 0, phase = 0, qf_phase = 0, acc = 0, acc_phase = 0;
    int cur_img = -1;
    int img = img_begin, tt = tt_begin;
    for (int t = t_begin; t < t_end; ++t) {
      if (img != cur_img) { mbar_wait(bar_qfull, qf_phase); qf_phase ^= 1; cur_img = img; }
      mbar_wait(bar_accempty + 8 * acc, acc_phase ^ 1);  // epilogue has drained this accumulator stage
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * kMaxNB;
      for (int ks = 0; ks < num_slabs; ++ks) {
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        const uint64_t da = smem_desc_sw128(smem_a + stage * kAStageBytes);
        const uint64_t db = smem_desc_sw128(smem_q + ks * q_slab_bytes);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kSlabCh / 8; ++kk)  // 8 tf32 = 32 bytes along K per instruction: +2 in 16-byte units
            umma_tf32(tmem_d, da + 2 * kk, db + 2 * kk, idesc, (ks | kk) ? 1u : 0u);
          umma_commit(bar_empty + 8 * stage);  // frees the memory-tile stage when those MMAs retire
        }
        __syncwarp();
        if (++stage == num_stages) { stage = 0; phase ^= 1; }
      }
      const bool last_of_image = (t + 1 == t_end) || (tt + 1 == p.tiles_per_image);
      if (elect_one()) {
        umma_commit(bar_accfull + 8 * acc);          // accumulator ready for the epilogue
        if (last_of_image) umma_commit(bar_qempty);  // the resident query block may be replaced
      }
      __syncwarp();
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
      if (++tt == p.tiles_per_image) { tt = 0; ++img; }
    }
  } else {
    // ================================================================ epilogue: thread = token (TMEM lane)
    // Two sets of four warps alternate over the tiles, so each SM sub-partition always has two epilogue warps to
    // interleave (a single warp per sub-partition is bound by its own dependent-issue latency).
    const int sub = warp & 3;               // TMEM sub-partition this warp may read: lanes [32*sub, 32*sub + 32)
    const int eset = (warp - 2) >> 2;
    const bool single = p.nblk == 1;
    const float scale = p.score_scale;
    // this set takes every kEpiSets-th tile; (acc, acc_phase) and (img, tt) follow the tile ordinal incrementally
    uint32_t acc = (uint32_t)eset % kAccStages, acc_phase = (uint32_t)eset / kAccStages;
    int img = img_begin, tt = tt_begin + eset;
    while (tt >= p.tiles_per_image) { tt -= p.tiles_per_image; ++img; }
    for (int t = t_begin + eset; t < t_end; t += kEpiSets) {
      const int K = p.box_start[img + 1] - p.box_start[img];
      const int kv = max(0, min(p.NB, K - b * p.NB));  // valid query columns of this block
      const float4* __restrict__ cj = reinterpret_cast<const float4*>(p.cpad + ((long long)img * p.nblk + b) * p.NB);
      mbar_wait(bar_accfull + 8 * acc, acc_phase);
      tc_fence_after();
      // the null logit 0 is part of the softmax; with several query blocks it is added once, by the combine kernel
      float mx = single ? 0.f : -1e30f, den0 = single ? 1.f : 0.f, den1 = 0.f, num0 = 0.f, num1 = 0.f;
      const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + acc * kMaxNB;
      for (int c0 = 0; c0 < kv; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        tmem_ld_wait();
        softmax_chunk(v, cj + (c0 >> 2), kv - c0, scale, mx, den0, den1, num0, num1);
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
      acc += kEpiSets;
      if (acc >= kAccStages) { acc -= kAccStages; acc_phase ^= 1; }
      tt += kEpiSets;
      while (tt >= p.tiles_per_image) { tt -= p.tiles_per_image; ++img; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ CTA-pair kernel
// More than kMaxNB queries per image: two CTAs of a cluster (one TPC) work as one tcgen05.mma.cta_group::2 unit on a
// 256-token tile.  Each CTA streams ITS 128 tokens of the memory (every byte is read once, by one SM) and keeps HALF of
// the query block resident (<= 160 rows), so up to 320 queries are contracted per byte of memory read -- twice the
// arithmetic intensity of the single-CTA kernel at the same shared-memory footprint.  The leader CTA issues the MMAs
// (two per k-step when NB > 256: N = n_lo and N = n_hi); completion is multicast to the barriers of both CTAs.
//  * full barriers live in the leader and take ONE arrival, the leader's expect_tx of the bytes of both CTAs; the
//    peer's TMA only completes bytes on it (a peer arrival per stage cost a cluster-scope release each: -25 % time);
//  * TMEM: NS = 512 / n_lo rotating accumulator slots, one per MMA part, so the epilogue of a tile drains its slots
//    while the MMAs of the next tile fill others; a part is read into registers first and its slot handed back before
//    the softmax runs;
//  * the two epilogue warp sets split the 32-column chunks of a part and merge their (max, den, num) through shared
//    memory, so with one query block the CTA writes final weights and no combine kernel runs.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kQmemThreads, 1)
qmem_weight_pair_kernel(const __grid_constant__ CUtensorMap tmap_mem, const __grid_constant__ CUtensorMap tmap_q,
                        const __grid_constant__ QmemParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int num_slabs = p.C / kSlabCh;
  const int half = p.NB / 2;                                // query rows resident in this CTA
  const uint32_t q_slab_bytes = (uint32_t)half * 128u;
  const uint32_t smem_q = smem_base;
  const uint32_t smem_a = smem_q + (uint32_t)num_slabs * q_slab_bytes;
  const int num_stages = p.stages;
  const uint32_t bars = smem_a + num_stages * kAStageBytes;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * kMaxAStages, bar_qfull = bars + 16 * kMaxAStages,
                 bar_qempty = bar_qfull + 8, bar_accfull = bar_qfull + 16, bar_accempty = bar_accfull + 8 * kPairSlots,
                 tmem_slot = bar_accempty + 8 * kPairSlots, xch = bars + 512;  // xch: 2 x 3 x 128 floats
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // warp index, provably uniform
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  // accumulator slots: every MMA part of a tile (one for NB <= 256, two above) takes the next of NS slots of SW columns,
  // so the epilogue of a tile drains its slots while the MMAs of the next tile fill others
  const int parts = p.n_hi ? 2 : 1;
  const int SW = p.n_lo;                                   // n_hi <= n_lo
  const int NS = min(kPairSlots, kTmemCols / SW);          // 2 (SW <= 256), 3 (SW = 160), 4 (SW <= 128)

  // work of this pair: query block b of a contiguous range of (image, 256-token tile) pairs
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int b = pair_id % p.nblk;
  const long long g = pair_id / p.nblk, G = num_pairs / p.nblk;
  const int t_begin = (int)(g * p.total_tiles / G), t_end = (int)((g + 1) * p.total_tiles / G);
  const int img_begin = t_begin / p.tiles_per_image, tt_begin = t_begin % p.tiles_per_image;

  if (threadIdx.x == 0) {
    // full barriers: ONE arrival (the leader's expect_tx of the bytes of both CTAs); the peer's loads only complete_tx
    for (int s = 0; s < num_stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_qfull, 1);
    mbar_init(bar_qempty, 1);
    for (int sl = 0; sl < kPairSlots; ++sl) {
      mbar_init(bar_accfull + 8 * sl, 1);
      mbar_init(bar_accempty + 8 * sl, 2 * 4 * kEpiSets);  // one arrival per epilogue warp of both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);

  if (warp == 0) {
    // ================================================================ TMA producer (both CTAs; signals the leader's barriers)
    uint32_t stage = 0, phase = 0, qe_phase = 0;
    int cur_img = -1;
    int img = img_begin, tt = tt_begin;
    for (int t = t_begin; t < t_end; ++t) {
      if (img != cur_img) {
        if (cur_img >= 0) { mbar_wait(bar_qempty, qe_phase); qe_phase ^= 1; }
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(bar_qfull, 2u * (uint32_t)num_slabs * q_slab_bytes);
          for (int ks = 0; ks < num_slabs; ++ks)
            tma_load_2d_pair(&tmap_q, bar_qfull, smem_q + ks * q_slab_bytes, ks * kSlabCh,
                             ((img * p.nblk + b) * 2 + (int)rank) * half);
        }
        __syncwarp();
        cur_img = img;
      }
      for (int ks = 0; ks < num_slabs; ++ks) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(bar_full + 8 * stage, 2u * kAStageBytes);
          tma_load_3d_pair(&tmap_mem, bar_full + 8 * stage, smem_a + stage * kAStageBytes, ks * kSlabCh, img,
                           tt * 2 * kTokTile + (int)rank * kTokTile);
        }
        __syncwarp();
        if (++stage == num_stages) { stage = 0; phase ^= 1; }
      }
      if (++tt == p.tiles_per_image) { tt = 0; ++img; }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (leader CTA only; whole warp, one lane issues)
    if (leader) {
      const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(2 * kTokTile >> 4) << 24);
      const uint32_t idesc_lo = idesc_base | ((uint32_t)(p.n_lo >> 3) << 17);
      const uint32_t idesc_hi = idesc_base | ((uint32_t)(p.n_hi >> 3) << 17);
      const uint32_t hi_row_bytes = (uint32_t)(p.n_lo / 2) * 128u;  // rows of the second MMA inside each resident slab
      uint32_t stage = 0, phase = 0, qf_phase = 0;
      // MMA parts take the accumulator slots round robin: (slot, round parity) of the next part
      uint32_t slot = 0, round = 0;
      int cur_img = -1, img = img_begin, tt = tt_begin;
      for (int t = t_begin; t < t_end; ++t) {
        if (img != cur_img) { mbar_wait(bar_qfull, qf_phase); qf_phase ^= 1; cur_img = img; }
        const uint32_t s_lo = slot;
        mbar_wait(bar_accempty + 8 * s_lo, round ^ 1);       // the epilogues have drained the slot
        if (++slot == (uint32_t)NS) { slot = 0; round ^= 1; }
        const uint32_t s_hi = slot;
        if (parts == 2) {
          mbar_wait(bar_accempty + 8 * s_hi, round ^ 1);
          if (++slot == (uint32_t)NS) { slot = 0; round ^= 1; }
        }
        tc_fence_after();
        const uint32_t d_lo = tmem_base + s_lo * SW, d_hi = tmem_base + s_hi * SW;
        for (int ks = 0; ks < num_slabs; ++ks) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint64_t da = smem_desc_sw128(smem_a + stage * kAStageBytes);
          const uint64_t db = smem_desc_sw128(smem_q + ks * q_slab_bytes);
          const uint64_t dbh = smem_desc_sw128(smem_q + ks * q_slab_bytes + hi_row_bytes);
          if (elect_one()) {
            if (parts == 2) {
#pragma unroll
              for (int kk = 0; kk < kSlabCh / 8; ++kk) {
                umma_tf32_pair(d_lo, da + 2 * kk, db + 2 * kk, idesc_lo, (ks | kk) ? 1u : 0u);
                umma_tf32_pair(d_hi, da + 2 * kk, dbh + 2 * kk, idesc_hi, (ks | kk) ? 1u : 0u);
              }
            } else {
#pragma unroll
              for (int kk = 0; kk < kSlabCh / 8; ++kk) umma_tf32_pair(d_lo, da + 2 * kk, db + 2 * kk, idesc_lo, (ks | kk) ? 1u : 0u);
            }
            umma_commit_pair(bar_empty + 8 * stage);
          }
          __syncwarp();
          if (++stage == num_stages) { stage = 0; phase ^= 1; }
        }
        const bool last_of_image = (t + 1 == t_end) || (tt + 1 == p.tiles_per_image);
        if (elect_one()) {
          umma_commit_pair(bar_accfull + 8 * s_lo);
          if (parts == 2) umma_commit_pair(bar_accfull + 8 * s_hi);
          if (last_of_image) umma_commit_pair(bar_qempty);
        }
        __syncwarp();
        if (++tt == p.tiles_per_image) { tt = 0; ++img; }
      }
      // the peer's last remote arrivals (on the last min(parts issued, NS) slots) have landed before this CTA may leave
      const int issued = (t_end - t_begin) * parts;
      for (int k = 0; k < min(issued, NS); ++k) {
        if (slot == 0) { slot = (uint32_t)NS; round ^= 1; }
        --slot;
        mbar_wait(bar_accempty + 8 * slot, round);
      }
    }
  } else {
    // ================================================================ epilogue (both CTAs): thread = token
    // The two warp sets split the 32-column chunks of every part between them.  A part of up to 192 columns is first
    // drained into registers (at most three chunks per set), its slot is handed back to the MMA issuer at once, and the
    // softmax runs on the registers while the tensor cores already fill the slot again.  The sets then merge their
    // (max, den, num) through shared memory, so one result per token leaves the CTA.
    const int sub = warp & 3;
    const int eset = (warp - 2) >> 2;
    const bool single = p.nblk == 1;
    const float scale = p.score_scale;
    const int row = sub * 32 + lane;
    uint32_t slot = 0, round = 0;  // accumulator slot and round parity of the next MMA part, as in the issuer
    int img = img_begin, tt = tt_begin;
    for (int t = t_begin; t < t_end; ++t) {
      const int K = p.box_start[img + 1] - p.box_start[img];
      const int kv = max(0, min(p.NB, K - b * p.NB));
      const float4* __restrict__ cj = reinterpret_cast<const float4*>(p.cpad + ((long long)img * p.nblk + b) * p.NB);
      float mx = -1e30f, den0 = 0.f, den1 = 0.f, num0 = 0.f, num1 = 0.f;
      for (int part = 0; part < parts; ++part) {
        const uint32_t sl = slot;
        mbar_wait(bar_accfull + 8 * sl, round);
        if (++slot == (uint32_t)NS) { slot = 0; round ^= 1; }
        tc_fence_after();
        const int q0 = part * p.n_lo;                         // first query (= cpad column) of this part
        const int kvp = max(0, min(part ? p.n_hi : p.n_lo, kv - q0));
        const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + sl * SW;
        const int ca = 32 * eset, cb = ca + 32 * kEpiSets, cc = cb + 32 * kEpiSets;
        if (kvp <= 3 * 32 * kEpiSets) {
          uint32_t va[32], vb[32], vc[32];
          if (ca < kvp) tmem_ld32(taddr + ca, va);
          if (cb < kvp) tmem_ld32(taddr + cb, vb);
          if (cc < kvp) tmem_ld32(taddr + cc, vc);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(bar_accempty + 8 * sl, 0);
          if (ca < kvp) softmax_chunk(va, cj + ((q0 + ca) >> 2), kvp - ca, scale, mx, den0, den1, num0, num1);
          if (cb < kvp) softmax_chunk(vb, cj + ((q0 + cb) >> 2), kvp - cb, scale, mx, den0, den1, num0, num1);
          if (cc < kvp) softmax_chunk(vc, cj + ((q0 + cc) >> 2), kvp - cc, scale, mx, den0, den1, num0, num1);
        } else {
          for (int c0 = ca; c0 < kvp; c0 += 32 * kEpiSets) {
            uint32_t v[32];
            tmem_ld32(taddr + c0, v);
            tmem_ld_wait();
            softmax_chunk(v, cj + ((q0 + c0) >> 2), kvp - c0, scale, mx, den0, den1, num0, num1);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(bar_accempty + 8 * sl, 0);
        }
      }
      // merge the two warp sets: set 1 publishes, set 0 combines and stores (double-buffered by tile parity)
      float den = den0 + den1, num = num0 + num1;
      float* x = reinterpret_cast<float*>(smem_raw + (xch - smem_u32(smem_raw))) + (size_t)(t & 1) * 3 * kTokTile;
      if (eset == 1) { x[row] = mx; x[kTokTile + row] = den; x[2 * kTokTile + row] = num; }
      asm volatile("bar.sync 1, %0;" ::"n"(128 * kEpiSets) : "memory");
      if (eset == 0) {
        const float mo = x[row], dn = x[kTokTile + row], nu = x[2 * kTokTile + row];
        // with one query block the null logit 0 joins here and the weight is final; otherwise qmem_combine_kernel adds it
        const float m = fmaxf(fmaxf(mx, mo), single ? 0.f : -1e30f);
        const float ra = ex2_approx(mx - m), rb = ex2_approx(mo - m);
        den = fmaf(den, ra, dn * rb) + (single ? ex2_approx(-m) : 0.f);
        num = fmaf(num, ra, nu * rb);
        const int tok = tt * 2 * kTokTile + (int)rank * kTokTile + row;
        if (tok < p.S) {
          if (single) {
            p.weight[(long long)img * p.S + tok] = sqrtf(__fdividef(num, den));
          } else {
            float* o = p.part + ((long long)img * p.nblk + b) * 3 * p.S + tok;
            o[0] = m;
            o[p.S] = den;
            o[2 * (long long)p.S] = num;
          }
        }
      }
      if (++tt == p.tiles_per_image) { tt = 0; ++img; }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// Query rows of every (image, block), zero padded to NB rows: Qpad[N, nblk, NB, C] and cpad[N, nblk, NB] (cpad in
// accumulator-column order).  Pair mode: the first NB/2 rows of a block are resident in the leader CTA, the rest in its
// peer, and accumulator column j of the MMA with N = n (n_lo, then n_hi) comes from row j of CTA 0 for j < n/2 and from
// row j - n/2 of CTA 1 otherwise -- the rows are laid out so that the accumulator columns are the queries in order.
__global__ void __launch_bounds__(256) qmem_gather_kernel(const float* __restrict__ hs_teacher, const int64_t* __restrict__ keepid,
                                                          const float* __restrict__ scores, const int* __restrict__ box_start,
                                                          int nblk, int NB, int pair, int n_lo, int n_hi, int C, int total_rows,
                                                          int64_t num_rows, float* __restrict__ qpad, float* __restrict__ cpad) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;  // warp = row (img * nblk + b) * NB + rr
  if (r >= total_rows) return;
  const int rr = r % NB, ib = r / NB, b = ib % nblk, img = ib / nblk;
  int col = rr;
  if (pair) {
    const int half = NB / 2, cta = rr / half, lr = rr % half, hl = n_lo / 2;
    col = lr < hl ? cta * hl + lr : n_lo + cta * (n_hi / 2) + (lr - hl);
  }
  const int first = __ldg(box_start + img), q = b * NB + col, K = __ldg(box_start + img + 1) - first;
  const bool valid = q < K;
  int64_t src = 0;
  if (valid) {
    src = __ldg(keepid + first + q);
    if (src < 0 || src >= num_rows) src = 0;  // defensive: never read outside hs_teacher
  }
  const float4* s4 = reinterpret_cast<const float4*>(hs_teacher + src * C);
  float4* d4 = reinterpret_cast<float4*>(qpad + (int64_t)r * C);
  for (int c = lane; c < C / 4; c += 32) d4[c] = valid ? __ldg(s4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  if (lane == 0) cpad[(int64_t)ib * NB + col] = valid ? (scores ? __ldg(scores + first + q) : 1.f) : 0.f;
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
  bool pair;                 // CTA-pair kernel (cta_group::2): more than kMaxNB queries per image
  int nblk, NB, n_lo, n_hi;  // query blocks per image, padded queries per block, MMA widths (pair mode)
  int part_blocks;           // partial (max, den, num) planes per image; 1 = the kernel writes the weights itself
  int64_t qpad, cpad, part, total;
  QmemPlan(int N, int64_t S, int C, int kmax) {
    // More than kMaxNB queries per image: the CTA-pair kernel reads the memory once for up to 320 queries where the
    // single-CTA kernel reads it once per 160 (B200, 16 images: 103 vs 131 us at 300 queries, 242 vs 301 us at 900;
    // profiles/r2/qmem_sweep.txt).  DSKD_QMEM_MODE = "1" / "2" forces one kernel (tests, tools).
    const char* mode = getenv("DSKD_QMEM_MODE");
    pair = mode != nullptr && (mode[0] == '1' || mode[0] == '2') ? mode[0] == '2' : kmax > kMaxNB;
    const int cap = pair ? 2 * kMaxNB : kMaxNB;
    nblk = std::max(1, (kmax + cap - 1) / cap);
    const int per = (std::max(kmax, 1) + nblk - 1) / nblk;
    NB = std::max((per + 15) / 16 * 16, pair ? 32 : 16);
    n_lo = NB; n_hi = 0;
    if (pair && NB > 256) { n_lo = kMaxNB; n_hi = NB - kMaxNB; }
    part_blocks = nblk;
    int64_t off = 0;
    auto up = [](int64_t x) { return (x + 255) / 256 * 256; };
    qpad = off; off += up((int64_t)N * nblk * NB * C * 4);
    cpad = off; off += up((int64_t)N * nblk * NB * 4);
    part = off; off += (part_blocks > 1) ? up((int64_t)N * part_blocks * 3 * S * 4) : 0;
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
  float* part = plan.part_blocks > 1 ? reinterpret_cast<float*>(base + plan.part) : nullptr;
  const int rows = a->N * plan.nblk * plan.NB;
  qmem_gather_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, st>>>(a->d_hs_teacher, a->d_keepid, a->d_scores, a->d_box_start, plan.nblk,
                                                                  plan.NB, plan.pair ? 1 : 0, plan.n_lo, plan.n_hi, a->C, rows,
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
    const cuuint32_t box[2] = {kSlabCh, (cuuint32_t)(plan.pair ? plan.NB / 2 : plan.NB)};
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
  const int tile_tokens = plan.pair ? 2 * kTokTile : kTokTile;
  p.tiles_per_image = (int)ceil_div(a->S, tile_tokens);
  p.total_tiles = (long long)p.tiles_per_image * a->N;
  DSKD_REQUIRE(p.total_tiles < (1ll << 31), "dskd_qmem_cell_weights: too many token tiles (%lld)", p.total_tiles);
  p.score_scale = 1.4426950408889634f / (sqrtf((float)a->C) * a->temperature);
  p.box_start = a->d_box_start;
  p.cpad = cpad;
  p.part = part;
  p.weight = a->d_cell_weight;
  p.n_lo = plan.n_lo;
  p.n_hi = plan.n_hi;
  const int resident_rows = plan.pair ? plan.NB / 2 : plan.NB;
  const size_t kSmemMax = 227 * 1024;  // 1024 alignment slack, resident queries, 512 barriers (+ the pair kernel's merge buffer)
  const size_t fixed = 1024 + (size_t)resident_rows * a->C * 4 + 512 + (plan.pair ? 2 * 3 * kTokTile * sizeof(float) : 0);
  int stages = (int)std::min<size_t>(kMaxAStages, (kSmemMax - fixed) / kAStageBytes);
  DSKD_REQUIRE(stages >= 2, "dskd_qmem_cell_weights: not enough shared memory for the memory-tile ring");
  p.stages = stages;
  const size_t smem = fixed + (size_t)stages * kAStageBytes;
  const int units = plan.pair ? kNumSMs / 2 : kNumSMs;  // CTAs or CTA pairs that fit the chip, one per SM
  const long long groups = std::max(1ll, std::min<long long>(units / plan.nblk, p.total_tiles));
  if (plan.pair) {
    DSKD_CUDA_OK(cudaFuncSetAttribute(qmem_weight_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    qmem_weight_pair_kernel<<<(unsigned)(2 * groups * plan.nblk), kQmemThreads, smem, st>>>(tmap_mem, tmap_q, p);
    DSKD_LAUNCH_OK("qmem_weight_pair_kernel");
  } else {
    DSKD_CUDA_OK(cudaFuncSetAttribute(qmem_weight_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    qmem_weight_kernel<<<(unsigned)(groups * plan.nblk), kQmemThreads, smem, st>>>(tmap_mem, tmap_q, p);
    DSKD_LAUNCH_OK("qmem_weight_kernel");
  }
  if (plan.part_blocks > 1) {
    const int64_t total = (int64_t)a->N * a->S;
    qmem_combine_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(part, a->N, plan.part_blocks, (int)a->S, a->d_cell_weight);
    DSKD_LAUNCH_OK("qmem_combine_kernel");
  }
  return DSKD_OK;
}
