// hamming_tc.cu — brute-force ORB-256 Hamming kNN (k=2) on the 5th-generation tensor cores.
//
// Same contract as hamming.cu (cv::BFMatcher(NORM_HAMMING)::knnMatch(k=2) as called from
// LoopClosureDetector::computeMatchedIndices — SURVEY.md A.4, B.1;
// /root/reference/images/kimera-multi.drawio:2583-2586, 2638): per query descriptor the two
// nearest train descriptors as packed keys (dist << 20 | trainIdx), ties to the lowest trainIdx.
//
// tcgen05.mma has no 1-bit kind and mma.sync b1 is emulated on sm_100a, so the
// descriptors are expanded to int8 +-1 (bit 1 -> +1, bit 0 -> -1): for two 256-bit descriptors
//     sum_k a_k b_k = 256 - 2 * hamming(a, b)      =>   hamming = (256 - dot) / 2,
// exact in the s32 accumulator.  One CTA per job (query frame x train range), persistent over
// the job list:
//   * every thread expands packed descriptors into shared memory in the canonical K-major
//     SWIZZLE_128B operand layout (rows of 128 bytes, 16-byte chunk c of row r stored at chunk
//     c ^ (r & 7); two 128-byte K blocks per 256-element row);
//   * one thread issues tcgen05.mma.cta_group::1.kind::i8, M = 128 query rows x N = 256 train
//     rows x K = 32 per instruction, 8 instructions per tile, accumulators in TMEM (two 256-column
//     buffers, so the tensor pipe works on tile t+1 while the CUDA cores read tile t);
//   * the epilogue reads the accumulator back with tcgen05.ld (thread = query row, 32 columns per
//     load) and keeps best / second best per row on 16-bit keys, two to a register (tc_chunk16):
//     one IMAD per distance, five packed min / max per four distances; the 32-bit key
//     hamming << 20 | train index is formed once per (query tile, train tile, 64-column quarter).
// Bound: the CUDA-core side (2.25 instructions per distance + the expansion), DESIGN.md §5.2.
#include "common.cuh"
#include "kernels.h"

namespace kml {

#ifndef KML_HOST_EMULATION

constexpr int kTcThreads = 512;                   // 16 warps: four per TMEM lane quarter, each reads 64 of a tile's 256 columns
constexpr int kTcM = 128, kTcN = 256;             // one MMA tile: 128 query rows x 256 train rows
constexpr int kTcMTiles = 4;                      // query rows resident per pass: 512
constexpr uint32_t kTcASlab = kTcM * 128;         // one K block of an A tile: 16 KB
constexpr uint32_t kTcATile = 2 * kTcASlab;       // 32 KB
constexpr uint32_t kTcBSlab = kTcN * 128;         // one K block of the B tile: 32 KB
constexpr uint32_t kTcBTile = 2 * kTcBSlab;       // 64 KB
constexpr uint32_t kTcSmemA = kTcMTiles * kTcATile;            // 128 KB
constexpr uint32_t kTcSmemBytes = kTcSmemA + kTcBTile + 1024;  // + alignment slack

// ---- tcgen05 / TMEM wrappers (PTX ISA, sm_100a) -----------------------------------------------
__device__ __forceinline__ void tmem_alloc_512(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_512(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// writes of the generic proxy (st.shared) must be made visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// shared-memory operand descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// instruction descriptor: D = s32, A = B = signed int8, both K-major, N = 256, M = 128
constexpr uint32_t kTcIdesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, bool accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(kTcIdesc), "r"((uint32_t)accumulate)
      : "memory");
}
// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int32_t* v) {
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 4 descriptor bits -> 4 int8 lanes of +1 / -1 (bit k of n in byte k)
__device__ __forceinline__ uint32_t expand4(uint32_t n) {
  const uint32_t w01 = (n * 0x00204081u) & 0x01010101u;
  return w01 | ((w01 ^ 0x01010101u) * 0xFFu);
}

// Expansion of packed descriptors (32 B each) into an operand tile (two K-block slabs of rows x
// 128 B, slab_bytes apart).  A work item is (row r, K block kb): one 128-bit load -> the 128 bytes
// of row r in slab kb = eight 16-byte chunks, chunk c stored at chunk c ^ (r & 7) (SWIZZLE_128B).
// Items are numbered it = 2 r + kb; a quarter warp (8 consecutive items = 4 rows x 2 K blocks)
// stores chunk (j + 4 kb) & 7 in its j-th store, which lands on 8 distinct 16-byte bank groups.
__device__ __forceinline__ uint4 tc_load_item(const uint8_t* __restrict__ src, int it, int nrows) {
  return (it < 2 * nrows) ? __ldg(reinterpret_cast<const uint4*>(src) + it) : make_uint4(0u, 0u, 0u, 0u);
}
// The expansion itself is a table look-up: lut[d] = the eight +-1 bytes of descriptor byte d (2 KB of shared
// memory, filled by the CTA when it starts), two look-ups per 16-byte chunk instead of ~20 integer instructions.
__device__ __forceinline__ void tc_store_item(uint8_t* tile, uint32_t slab_bytes, int it, const uint4& x, const uint2* lut) {
  const int r = it >> 1, kb = it & 1;
  uint8_t* row = tile + (uint32_t)kb * slab_bytes + (uint32_t)r * 128u;
  // K block 1 starts with its upper four chunks (c = (j + 4 kb) & 7 = j ^ 4 kb): rotate the words once, then
  // store j holds chunk j of the rotated block and goes to chunk position j ^ s2
  const uint32_t w[4] = {kb ? x.z : x.x, kb ? x.w : x.y, kb ? x.x : x.z, kb ? x.y : x.w};
  const int s2 = (r & 7) ^ (4 * kb);
  const unsigned char* lb = reinterpret_cast<const unsigned char*>(lut);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t h = w[j >> 1];
    const uint32_t o0 = (j & 1) ? ((h >> 13) & 0x7F8u) : ((h << 3) & 0x7F8u);   // 8 * descriptor byte 2c
    const uint32_t o1 = (j & 1) ? ((h >> 21) & 0x7F8u) : ((h >> 5) & 0x7F8u);   // 8 * descriptor byte 2c + 1
    const uint2 e0 = *reinterpret_cast<const uint2*>(lb + o0), e1 = *reinterpret_cast<const uint2*>(lb + o1);
    *reinterpret_cast<uint4*>(row + ((j ^ s2) << 4)) = make_uint4(e0.x, e0.y, e1.x, e1.y);
  }
}
// rows [0, nrows) of `src` into `tile` (query side: no register prefetch needed, once per pass)
__device__ __forceinline__ void tc_expand_rows(uint8_t* tile, uint32_t slab_bytes, const uint8_t* __restrict__ src,
                                               int nrows, int tid, const uint2* lut) {
  for (int it0 = 0; it0 < 2 * nrows; it0 += 2 * kTcThreads) {
    const uint4 x0 = tc_load_item(src, it0 + tid, nrows);
    const uint4 x1 = tc_load_item(src, it0 + kTcThreads + tid, nrows);
    if (it0 + tid < 2 * nrows) tc_store_item(tile, slab_bytes, it0 + tid, x0, lut);
    if (it0 + kTcThreads + tid < 2 * nrows) tc_store_item(tile, slab_bytes, it0 + kTcThreads + tid, x1, lut);
  }
}

// The two smallest keys of a row.  Keys are distinct (the train index sits in the low bits), so the two
// smallest keys of a union of two (best, second) pairs are min(b0, b1) and min(max(b0, b1), min(s0, s1)).
// The epilogue works on 16-bit keys two to a register (VIMNMX.U16x2 / VIMNMX3.U16x2).  Inside one (query tile,
// train tile, 64-column quarter) a key fits 16 bits: ham << 7 | local column, ham = (256 - dot) / 2 <= 256,
// i.e. -64 dot + 16384 + column.  Register lane 0 follows the even columns, lane 1 the odd ones; two
// registers (four columns) are ordered against each other first, so four keys cost five min / max
// instructions instead of twelve, plus one IMAD each.  Columns past the valid ones become 0xFFFF.
template <int C, bool FULL>
__device__ __forceinline__ void tc_chunk16(const int32_t* v, int nvalid, uint32_t* pb, uint32_t* ps) {
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const int i0 = 4 * g;
    uint32_t P[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = i0 + 2 * h;
      const uint32_t k0 = 16384u + (uint32_t)(C * 32 + i), k1 = 16384u + (uint32_t)(C * 32 + i + 1);
      uint32_t x = (uint32_t)(v[i] * -64) + (k0 | (k1 << 16));
      x = (uint32_t)(v[i + 1] * -4194304) + x;
      if (!FULL) {
        if (i >= nvalid) x |= 0x0000FFFFu;
        if (i + 1 >= nvalid) x |= 0xFFFF0000u;
      }
      P[h] = x;
    }
    const uint32_t lo = __vminu2(P[0], P[1]), hi = __vmaxu2(P[0], P[1]);
    const uint32_t t = __vmaxu2(pb[g & 1], lo);
    pb[g & 1] = __vminu2(pb[g & 1], lo);
    ps[g & 1] = __vimin3_u16x2(t, ps[g & 1], hi);
  }
}
// 16-bit key of a region -> the kernel's 32-bit key ham << 20 | train index (0xFFFF -> no key)
__device__ __forceinline__ uint32_t tc_key32(uint32_t k16, uint32_t col0) {
  return k16 == 0xFFFFu ? 0xFFFFFFFFu : ((k16 >> 7) << 20) + (col0 + (k16 & 127u));
}
__device__ __forceinline__ void tc_merge2(uint32_t& b, uint32_t& s, uint32_t b1, uint32_t s1) {
  const uint32_t nb = min(b, b1);
  s = min(max(b, b1), min(s, s1));
  b = nb;
}

__global__ void __launch_bounds__(kTcThreads, 1) hamming_tc_kernel(const HamJob* __restrict__ jobs, int njobs) {
  extern __shared__ uint8_t tc_smem_raw[];
  __shared__ __align__(8) uint64_t full[2];
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t s_best[3][kTcMTiles * kTcM], s_second[3][kTcMTiles * kTcM];
  __shared__ uint2 s_lut[256];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + kTcSmemA;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp >> 2;                   // which 64 columns of a tile this warp reads
  const int row_in_tile = (warp & 3) * 32 + lane;  // TMEM lane == query row of the tile
  s_lut[threadIdx.x & 255] = make_uint2(expand4(threadIdx.x & 15u), expand4((threadIdx.x >> 4) & 15u));
  if (warp == 0) tmem_alloc_512(&tmem_slot);
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
  uint32_t ph[2] = {0u, 0u};  // phase parity of full[0], full[1] (every thread tracks both)

  // A CTA takes a CONTIGUOUS range of the job list: the candidate pairs of one query frame are
  // consecutive jobs with the same query side, whose expansion is then reused (it is a fifth of
  // the CUDA-core work of a pair).
  const int j_lo = (int)(((long long)njobs * blockIdx.x) / gridDim.x);
  const int j_hi = (int)(((long long)njobs * (blockIdx.x + 1)) / gridDim.x);
  const uint8_t* a_src = nullptr;  // what sA currently holds: rows [a_q0, a_q0 + a_nq) of a_src
  int a_q0 = -1, a_nq = -1;
  for (int job_i = j_lo; job_i < j_hi; ++job_i) {
    const HamJob job = jobs[job_i];
    if (job.nq <= 0) continue;
    for (int q0 = 0; q0 < job.nq; q0 += kTcMTiles * kTcM) {
      const int nq = min(kTcMTiles * kTcM, job.nq - q0);
      const int mtiles = (nq + kTcM - 1) / kTcM;
      uint32_t best[kTcMTiles], second[kTcMTiles];
#pragma unroll
      for (int m = 0; m < kTcMTiles; ++m) best[m] = second[m] = 0xFFFFFFFFu;
      // query side: up to 512 rows, tile m at sA + m * 32 KB (kept if the previous job left the same rows there)
      if (!(job.q == a_src && q0 == a_q0 && nq == a_nq)) {
        for (int m = 0; m < mtiles; ++m)
          tc_expand_rows(sA + m * kTcATile, kTcASlab, job.q + (size_t)(q0 + m * kTcM) * 32, min(kTcM, nq - m * kTcM), tid, s_lut);
        a_src = job.q; a_q0 = q0; a_nq = nq;
      }
      // train side: one 256-row tile at a time; the packed rows of the NEXT tile are loaded into
      // registers (2 items per thread) while the tensor pipe and the epilogue work on this one
      static_assert(kTcThreads == 2 * kTcN, "one (row, K block) item of the train tile per thread");
      uint4 pf0 = tc_load_item(job.t, tid, min(kTcN, job.nt));
      for (int n0 = 0; n0 < job.nt; n0 += kTcN) {
        const int nn = min(kTcN, job.nt - n0);
        if (tid < 2 * nn) tc_store_item(sB, kTcBSlab, tid, pf0, s_lut);
        if (n0 + kTcN < job.nt) {
          const int nn1 = min(kTcN, job.nt - n0 - kTcN);
          pf0 = tc_load_item(job.t + (size_t)(n0 + kTcN) * 32, tid, nn1);
        }
        fence_proxy_async();
        __syncthreads();
        auto issue = [&](int m) {
          tc_fence_after();
          const uint32_t d = tmem + (uint32_t)(m & 1) * kTcN;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t ko = (uint32_t)(k >> 2), ki = (uint32_t)(k & 3) * 32u;
            tc_mma_i8(d, tc_desc(sA_u + (uint32_t)m * kTcATile + ko * kTcASlab + ki),
                      tc_desc(sB_u + ko * kTcBSlab + ki), k > 0);
          }
          tc_commit(&full[m & 1]);
        };
        if (tid == 0) {
          issue(0);
          if (mtiles > 1) issue(1);
        }
        // chunks of 32 columns this thread reads in this tile: [quarter*64, quarter*64 + 64) cut at nn
        const int jl0 = quarter * 64;
        const int nch = nn > jl0 ? min(2, (nn - jl0 + 31) >> 5) : 0;
        for (int m = 0; m < mtiles; ++m) {
          const int buf = m & 1;
          mbar_wait(&full[buf], ph[buf]);
          ph[buf] ^= 1u;
          tc_fence_after();
          const uint32_t tbase = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(buf * kTcN + jl0);
          // this thread's region: row `row_in_tile` of tile m x the <= 64 columns [jl0, jl0 + 64) of the train tile
          // (one register buffer: with sixteen warps the other warps cover the TMEM load's latency)
          uint32_t pb[2] = {0xFFFFFFFFu, 0xFFFFFFFFu}, ps[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
          int32_t va[32];
          const int nloc = nn - jl0;  // valid columns of the region (may exceed 64)
          if (nch > 0) {
            tmem_ld32(tbase, va);
            tmem_ld_wait();                                        // va = chunk 0
            if (nloc >= 32) tc_chunk16<0, true>(va, 32, pb, ps); else tc_chunk16<0, false>(va, nloc, pb, ps);
          }
          if (nch > 1) {
            tmem_ld32(tbase + 32u, va);
            tmem_ld_wait();                                        // va = chunk 1
            if (nloc >= 64) tc_chunk16<1, true>(va, 32, pb, ps); else tc_chunk16<1, false>(va, nloc - 32, pb, ps);
          }
          // the region's two smallest keys as 32-bit keys, merged into the row's running pair
          const uint32_t col0 = (uint32_t)(n0 + jl0);
          uint32_t lb = tc_key32(pb[0] & 0xFFFFu, col0), ls = tc_key32(ps[0] & 0xFFFFu, col0);
          tc_merge2(lb, ls, tc_key32(pb[0] >> 16, col0), tc_key32(ps[0] >> 16, col0));
          tc_merge2(lb, ls, tc_key32(pb[1] & 0xFFFFu, col0), tc_key32(ps[1] & 0xFFFFu, col0));
          tc_merge2(lb, ls, tc_key32(pb[1] >> 16, col0), tc_key32(ps[1] >> 16, col0));
#pragma unroll
          for (int mm = 0; mm < kTcMTiles; ++mm)
            if (mm == m) tc_merge2(best[mm], second[mm], lb, ls);
          tc_fence_before();
          __syncthreads();  // every warp is done reading TMEM buffer `buf`
          if (tid == 0 && m + 2 < mtiles) issue(m + 2);
        }
        // all MMAs of this train tile have completed (their commits were waited for): sB is free
      }
      // merge the four column quarters of every row and write the keys
      if (quarter > 0) {
#pragma unroll
        for (int m = 0; m < kTcMTiles; ++m) {
          s_best[quarter - 1][m * kTcM + row_in_tile] = best[m];
          s_second[quarter - 1][m * kTcM + row_in_tile] = second[m];
        }
      }
      __syncthreads();
      if (quarter == 0) {
#pragma unroll
        for (int m = 0; m < kTcMTiles; ++m) {
          const int qi = m * kTcM + row_in_tile;
          if (qi < nq) {
            uint32_t b = best[m], s = second[m];
#pragma unroll
            for (int o = 0; o < 3; ++o) tc_merge2(b, s, s_best[o][qi], s_second[o][qi]);
            job.keys[2 * (size_t)(q0 + qi) + 0] = b;
            job.keys[2 * (size_t)(q0 + qi) + 1] = s;
          }
        }
      }
      __syncthreads();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc_512(tmem);
}

void launch_hamming_jobs_tc(const HamJob* d_jobs, int njobs, cudaStream_t s) {
  if (njobs <= 0) return;
  KML_CUDA(cudaFuncSetAttribute(hamming_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
  const int grid = min(njobs, kNumSMs);
  hamming_tc_kernel<<<grid, kTcThreads, kTcSmemBytes, s>>>(d_jobs, njobs);
}

#else  // the SIMT emulator has no tensor cores: the POPC kernel computes the same keys

void launch_hamming_jobs_tc(const HamJob* d_jobs, int njobs, cudaStream_t s) { launch_hamming_jobs(d_jobs, njobs, 0, s); }

#endif

}  // namespace kml
