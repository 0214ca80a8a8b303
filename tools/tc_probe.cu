// tc_probe: hardware probe for the building blocks of csrc/mlp_tc.cu on sm_100a (run on the GPU box):
//   * 2-CTA cluster, tcgen05.alloc.cta_group::2, TMA tensor-map load with .cta_group::2 signalling the leader's mbarrier
//   * tcgen05.mma.cta_group::2.kind::f16 (M=128 over the pair, N=256, K=16) with no-swizzle K-major "panel" layouts
//     (panel = 8 K-elements x rows x 16 B; LBO = panel stride, SBO = 128 B)
//   * tcgen05.commit multicast, tcgen05.ld 32x32b.x32 read-back
// Operands are small integers so every product is exact: D[m][n] = 256*m + n.  The host decodes where each (m,n)
// landed in TMEM (cta, lane, column) and prints the layout, then times a chain of MMAs.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tools/tc_probe.cu
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int KT = 32;             // K extent of the probe tile (2 MMA k-steps)
constexpr int NPAN = KT / 8;       // K-group panels
constexpr int A_BYTES = NPAN * 64 * 16;    // 64 rows per CTA
constexpr int B_BYTES = NPAN * 128 * 16;   // 128 N-rows per CTA
constexpr int OFF_A = 0, OFF_B = 4096, OFF_BAR = 229376;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // layout_type 0 = SWIZZLE_NONE, base_offset 0
}
__device__ __forceinline__ void mma_f16_2cta(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void commit_mc(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmapB, float* out, long long* timing, int lbo_swap, int n_chain, int big_smem) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t rank = cluster_ctarank();
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_full = sbase + OFF_BAR, bar_done = sbase + OFF_BAR + 8, bar_done2 = sbase + OFF_BAR + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 32);

  if (tid == 0) {
    mbar_init(bar_full, 1);
    mbar_init(bar_done, 1);
    mbar_init(bar_done2, 1);
    mbar_init(sbase + OFF_BAR + 24, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  // A panels: row r (global m = 64*rank + r): A[m][0] = m, A[m][1] = 1, A[m][8] = 1, everything else 0
  {
    __half* A = reinterpret_cast<__half*>(smem + OFF_A);
    for (int i = tid; i < NPAN * 64 * 8; i += 128) {
      const int g = i / (64 * 8), r = (i / 8) % 64, e = i % 8, k = g * 8 + e;
      const int m = 64 * rank + r;
      float v = 0.f;
      if (k == 0) v = (float)m;
      if (k == 8) v = 1.f;
      A[i] = __float2half(v);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (tid == 0) {
    if (rank == 0) mbar_expect_tx(bar_full, 2 * B_BYTES);
    // TMA: this CTA's half of B -> own smem, complete_tx on the LEADER's barrier (peer bit cleared)
    const uint32_t bar_leader = bar_full & 0xFEFFFFFFu;
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            sbase + OFF_B),
        "l"(&tmapB), "r"(bar_leader), "r"(0), "r"((int)(rank * (B_BYTES / 128)))
        : "memory");
  }
  if (rank == 0 && tid == 0) {
    mbar_wait(bar_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_lbo = lbo_swap ? 128 : 64 * 16, a_sbo = lbo_swap ? 64 * 16 : 128;
    const uint32_t b_lbo = lbo_swap ? 128 : 128 * 16, b_sbo = lbo_swap ? 128 * 16 : 128;
    for (int ks = 0; ks < KT / 16; ++ks) {
      const uint64_t da = make_desc(sbase + OFF_A + ks * 2 * 64 * 16, a_lbo, a_sbo);
      const uint64_t db = make_desc(sbase + OFF_B + ks * 2 * 128 * 16, b_lbo, b_sbo);
      mma_f16_2cta(tmem, da, db, idesc, ks > 0 ? 1u : 0u);
    }
    commit_mc(bar_done);
  }
  mbar_wait(bar_done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // raw dump: out[rank][lane][col], 128 lanes x 256 columns
  for (int c0 = 0; c0 < 256; c0 += 32) {
    uint32_t r[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
        "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[((size_t)rank * 128 + tid) * 256 + c0 + j] = __uint_as_float(r[j]);
  }
  // ---- throughput: a chain of MMAs on the same operands
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (rank == 0 && tid == 0 && n_chain > 0) {
    const uint32_t idesc = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t da = make_desc(sbase + OFF_A, 64 * 16, 128);
    const uint64_t db = make_desc(sbase + OFF_B, 128 * 16, 128);
    const long long t0 = clock64();
    for (int i = 0; i < n_chain; ++i) mma_f16_2cta(tmem + 256, da, db, idesc, 1u);
    commit_mc(bar_done2);
    const long long t1 = clock64();
    mbar_wait(bar_done2, 0);
    const long long t2 = clock64();
    timing[0] = t1 - t0;
    timing[1] = t2 - t0;
  }
  if (n_chain > 0) mbar_wait(bar_done2, 0);
  // ---- realistic pattern: A 128 KB (hi 64 KB | lo 64 KB), W ring 6 x 16 KB at 128 KB, 3 passes per k16 step
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (rank == 0 && tid == 0 && n_chain > 0 && big_smem) {
    const uint32_t idesc = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t bar_done3 = sbase + OFF_BAR + 24;
    const long long t0 = clock64();
    int it = 0, cnt = 0;
    for (int layer = 0; layer < n_chain / 192; ++layer)
      for (int kc = 0; kc < 16; ++kc)
        for (int h = 0; h < 2; ++h, ++it) {
          const uint32_t wb = sbase + 131072 + (it % 6) * 16384;
          for (int ks = 0; ks < 2; ++ks) {
            const uint32_t a_off = (uint32_t)(kc * 4 + ks * 2) * 1024;
            const uint64_t a_hi = make_desc(sbase + a_off, 1024, 128), a_lo = make_desc(sbase + 65536 + a_off, 1024, 128);
            const uint64_t b_hi = make_desc(wb + ks * 4096, 2048, 128), b_lo = make_desc(wb + 8192 + ks * 4096, 2048, 128);
            mma_f16_2cta(tmem + h * 128, a_hi, b_hi, idesc, 1u);
            mma_f16_2cta(tmem + h * 128, a_lo, b_hi, idesc, 1u);
            mma_f16_2cta(tmem + h * 128, a_hi, b_lo, idesc, 1u);
            cnt += 3;
          }
        }
    commit_mc(bar_done3);
    mbar_wait(bar_done3, 0);
    const long long t2 = clock64();
    timing[2] = t2 - t0;
    timing[3] = cnt;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int n_chain = argc > 1 ? atoi(argv[1]) : 2048;
  // B blob: [cta half][panel g][128 n-rows][8 k] fp16;  B[n][0] = 256, B[n][1]... see below
  std::vector<__half> hB(2 * NPAN * 128 * 8);
  for (int h = 0; h < 2; ++h)
    for (int g = 0; g < NPAN; ++g)
      for (int r = 0; r < 128; ++r)
        for (int e = 0; e < 8; ++e) {
          const int k = g * 8 + e, n = h * 128 + r;
          float v = 0.f;
          if (k == 0) v = 256.f;
          if (k == 8) v = (float)n;
          hB[((size_t)(h * NPAN + g) * 128 + r) * 8 + e] = __float2half(v);
        }
  __half* dB;
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  float* dOut;
  CK(cudaMalloc(&dOut, 2 * 128 * 256 * 4));
  long long* dT;
  CK(cudaMalloc(&dT, 64));
  CK(cudaMemset(dT, 0, 64));

  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {64, (cuuint64_t)(hB.size() * 2 / 128)};
  const cuuint64_t gstr[1] = {128};
  const cuuint32_t box[2] = {64, (cuuint32_t)(B_BYTES / 128)};
  const cuuint32_t estr[2] = {1, 1};
  CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { printf("encode failed %d\n", (int)cr); return 1; }

  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 230400));
  for (int lbo_swap = 0; lbo_swap < 1; ++lbo_swap) {
    CK(cudaMemset(dOut, 0xFF, 2 * 128 * 256 * 4));
    probe_kernel<<<2, 128, 230400>>>(tmap, dOut, dT, lbo_swap, lbo_swap == 0 ? n_chain : 0, 1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("lbo_swap=%d: kernel failed: %s\n", lbo_swap, cudaGetErrorString(e)); return 1; }
    std::vector<float> o(2 * 128 * 256);
    CK(cudaMemcpy(o.data(), dOut, o.size() * 4, cudaMemcpyDeviceToHost));
    // decode: value = 256 m + n
    long ok_assumed = 0, decodable = 0;
    for (int c = 0; c < 2; ++c)
      for (int l = 0; l < 128; ++l)
        for (int col = 0; col < 256; ++col) {
          const float v = o[((size_t)c * 128 + l) * 256 + col];
          const int iv = (int)v;
          if (v == (float)iv && iv >= 0 && iv < 256 * 128) {
            ++decodable;
            const int m = iv / 256, n = iv % 256;
            // assumed "2x2" layout: cta = m/64, lane = (m%64) + 64*(n/128), col = n%128  (columns 128..255 unused)
            if (c == m / 64 && l == (m % 64) + 64 * (n / 128) && col == n % 128) ++ok_assumed;
          }
        }
    printf("lbo_swap=%d: decodable=%ld  matching-assumed-layout=%ld (expect 32768 in columns 0..127)\n", lbo_swap, decodable, ok_assumed);
    for (int c = 0; c < 2; ++c)
      for (int l : {0, 1, 63, 64, 65, 127}) {
        printf("  cta%d lane%3d:", c, l);
        for (int col : {0, 1, 2, 127, 128, 129, 255}) {
          const float v = o[((size_t)c * 128 + l) * 256 + col];
          const int iv = (int)v;
          if (v == (float)iv && iv >= 0 && iv < 32768) printf(" c%d=(m%d,n%d)", col, iv / 256, iv % 256);
          else printf(" c%d=%g", col, v);
        }
        printf("\n");
      }
    if (lbo_swap == 0) {
      long long t[4];
      CK(cudaMemcpy(t, dT, 32, cudaMemcpyDeviceToHost));
      printf("realistic operand pattern (A 128 KB, W ring 96 KB, 3 passes): %lld MMAs in %lld cyc -> %.1f cyc/MMA\n", t[3], t[2], t[3] ? (double)t[2] / t[3] : 0.0);
      printf("chain of %d MMAs (M=128 pair, N=256, K=16): issue %lld cyc, complete %lld cyc -> %.1f cyc/MMA\n", n_chain, t[0], t[1],
             (double)t[1] / n_chain);
    }
  }
  return 0;
}
