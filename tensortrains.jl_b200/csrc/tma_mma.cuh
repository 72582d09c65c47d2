// tma_mma.cuh -- sm_100a building blocks shared by the sweep and sampling kernels:
// 1-D TMA bulk copies (cp.async.bulk + mbarrier; SASS UBLKCP / SYNCS) and the FP64 tensor-core tile
// (mma.sync.aligned.m16n8k8.f64; SASS DMMA).  tcgen05/TMEM have no FP64 kind, so DMMA is the tensor path.
#pragma once
#include <stdint.h>

namespace ttb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy; bytes and both addresses must be multiples of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global bulk copy (bulk async-group completion); bytes and both addresses multiples of 16
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups are pending (their writes are then complete and visible)
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// order earlier generic-proxy accesses to shared memory before later async-proxy (TMA) writes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a transaction that never completes traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t it = 0; !mbar_try_wait(bar, parity); ++it)
    if (it > (1u << 26)) __trap();
}

// D(16x8) += A(16x8, row) * B(8x8, col), FP64.  Fragments (g = lane/4, t = lane%4):
//   a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4);   b0 (k=t, n=g)  b1 (k=t+4, n=g);
//   d0 (g, 2t) d1 (g, 2t+1) d2 (g+8, 2t) d3 (g+8, 2t+1)
__device__ __forceinline__ void dmma_16x8x8(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

// ---- 32 x 32 x 32 products of the small-bond kernels (env_small.cuh, twovar_mma.cuh, k_absorb_small) ----
// warp tile of a 32 x 32 x 32 product: rows 16 mt .. +16, columns 8 nt0 .. +16.
// A(m, k) at Aop[m * ARS + k * ACS], B(k, n) at Bop[k * BKS + n * BNS]
template <int ARS, int ACS, int BKS, int BNS>
__device__ __forceinline__ void mma32_warp(double (&d)[2][4], const double* Aop, const double* Bop, int mt, int nt0, int g,
                                           int tq) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const double* pa = Aop + (mt * 16 + g) * ARS + (ks * 8 + tq) * ACS;
    const double af[4] = {pa[0], pa[8 * ARS], pa[4 * ACS], pa[8 * ARS + 4 * ACS]};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const double* pb = Bop + (ks * 8 + tq) * BKS + ((nt0 + j) * 8 + g) * BNS;
      const double bf[2] = {pb[0], pb[4 * BKS]};
      dmma_16x8x8(d[j], af, bf);
    }
  }
}

// block max over 4 warps of NaN-propagating max-abs values (ordered bits: NaN > inf > finite)
__device__ __forceinline__ double block4_max_bits(double mx, double* red, int lane, int wid) {
  unsigned long long bits = (unsigned long long)__double_as_longlong(mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long ot = __shfl_xor_sync(0xffffffffu, bits, o);
    bits = ot > bits ? ot : bits;
  }
  if (lane == 0) red[wid] = __longlong_as_double((long long)bits);
  __syncthreads();
  unsigned long long bb = 0ull;
#pragma unroll
  for (int w2 = 0; w2 < 4; ++w2) {
    const unsigned long long v = (unsigned long long)__double_as_longlong(red[w2]);
    bb = v > bb ? v : bb;
  }
  return __longlong_as_double((long long)bb);
}

}  // namespace ttb

namespace ttb {
// ---- thread-block clusters: distributed shared memory ------------------------------------------------
// shared::cta address of THIS CTA -> shared::cluster address of the same location in CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// bulk copy from this CTA's shared memory into another CTA's; completion is counted on the REMOTE mbarrier
__device__ __forceinline__ void bulk_s2remote(uint32_t dst_cluster, const void* src, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(smem_u32(src)), "r"(bytes), "r"(bar_cluster)
               : "memory");
}
__device__ __forceinline__ void st_remote_release_s32(uint32_t addr_cluster, int v) {
  asm volatile("st.release.cluster.shared::cluster.s32 [%0], %1;" ::"r"(addr_cluster), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_cluster_s32(const int* p) {
  int v;
  asm volatile("ld.acquire.cluster.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
}  // namespace ttb
