// Bring-up kernel for the tcgen05 building blocks (tests only): one D[128 x Ncols] = A B^T tile with
// split operands through exactly the helpers the product kernels use (128B-swizzled K-major
// operand tiles, shared-memory matrix descriptors, instruction descriptor, TMEM load shapes).
#include "common.cuh"
#include "ptx.cuh"
#include "umma_tf32.cuh"

namespace sga {
namespace {

constexpr uint32_t kBlk = 16384;

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// kind 0: bf16 hi/lo (atom = 64 elements), kind 1: tf32 hi/lo (atom = 32 elements)
__global__ void __launch_bounds__(128, 1)
selftest_umma_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int Ncols,
                     int K, int kind) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sm_base = ptx::smem_u32(sm);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int atom = kind == 0 ? 64 : 32;        // elements per 128-byte row
  const int natom = K / atom;
  const int epc = kind == 0 ? 8 : 4;           // elements per 16-byte chunk
  // operand tiles: [A hi | A lo | B hi | B lo], each natom blocks of 16 KiB
  const uint32_t AHI = 0, ALO = natom * kBlk, BHI = 2 * natom * kBlk, BLO = 3 * natom * kBlk;
  for (int which = 0; which < 2; ++which) {
    const float* src = which ? B : A;
    const int rows = which ? Ncols : 128;
    const uint32_t HI = which ? BHI : AHI, LO = which ? BLO : ALO;
    const int chunks_per_row = K / epc;
    for (int i = tid; i < 128 * chunks_per_row; i += 128) {
      int r = i / chunks_per_row, j = i % chunks_per_row;
      int ka = (j * epc) / atom, jj = ((j * epc) % atom) / epc;
      uint32_t off = (uint32_t)ka * kBlk + ptx::sw128_offset(r, jj);
      uint4 hi = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
      if (r < rows) {
        const float* p = src + (int64_t)r * K + j * epc;
        if (kind == 0) {
          uint32_t h[4], l[4];
          for (int e = 0; e < 4; ++e) {
            float a = p[2 * e], b = p[2 * e + 1];
            h[e] = pack2(a, b);
            l[e] = pack2(a - __uint_as_float(h[e] << 16), b - __uint_as_float(h[e] & 0xFFFF0000u));
          }
          hi = make_uint4(h[0], h[1], h[2], h[3]);
          lo = make_uint4(l[0], l[1], l[2], l[3]);
        } else {
          uint32_t h[4], l[4];
          for (int e = 0; e < 4; ++e) {
            float a = p[e];
            h[e] = tf32x3::rn_tf32(a);
            l[e] = tf32x3::rn_tf32(a - __uint_as_float(h[e]));
          }
          hi = make_uint4(h[0], h[1], h[2], h[3]);
          lo = make_uint4(l[0], l[1], l[2], l[3]);
        }
      }
      *reinterpret_cast<uint4*>(sm + HI + off) = hi;
      *reinterpret_cast<uint4*>(sm + LO + off) = lo;
    }
  }
  if (tid == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) ptx::tmem_alloc<256>(&tmem_slot);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = ptx::make_idesc(kind == 0 ? 1 : 2, 128, Ncols);
    const int ksteps = atom / (kind == 0 ? 16 : 8);   // 4 k-steps of 32 bytes per atom
    bool first = true;
    for (int pass = 0; pass < 3; ++pass) {
      const uint32_t a_off = (pass == 1) ? ALO : AHI;
      const uint32_t b_off = (pass == 2) ? BLO : BHI;
      for (int ka = 0; ka < natom; ++ka)
        for (int ks = 0; ks < ksteps; ++ks) {
          uint64_t ad = ptx::smem_desc_sw128(sm_base + a_off + ka * kBlk + ks * 32);
          uint64_t bd = ptx::smem_desc_sw128(sm_base + b_off + ka * kBlk + ks * 32);
          if (kind == 0) ptx::umma_bf16(tmem, ad, bd, idesc, first ? 0u : 1u);
          else ptx::umma_tf32(tmem, ad, bd, idesc, first ? 0u : 1u);
          first = false;
        }
    }
    ptx::umma_commit(&bar);
  }
  ptx::mbar_wait(&bar, 0);
  ptx::tc_fence_after();
  const int row = 32 * warp + lane;
  for (int c0 = 0; c0 < Ncols; c0 += 16) {
    uint32_t v[16];
    ptx::tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16) + c0, v);
    ptx::tmem_ld_wait();
    for (int e = 0; e < 16; ++e) D[(int64_t)row * Ncols + c0 + e] = __uint_as_float(v[e]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<256>(tmem);
}

}  // namespace
}  // namespace sga

extern "C" int sga_selftest_umma(const float* A, const float* B, float* D, int Ncols, int K, int kind, void* stream) {
  SGA_REQUIRE(kind == 0 || kind == 1, "sga_selftest_umma: kind=%d", kind);
  const int atom = kind == 0 ? 64 : 32;
  SGA_REQUIRE(Ncols >= 16 && Ncols <= 128 && Ncols % 16 == 0, "sga_selftest_umma: Ncols=%d must be a multiple of 16 in 16..128", Ncols);
  SGA_REQUIRE(K >= atom && K % atom == 0 && K / atom <= 2, "sga_selftest_umma: K=%d must be 1 or 2 swizzle atoms of %d", K, atom);
  size_t smem = (size_t)4 * (K / atom) * sga::kBlk + 1024;
  static size_t attr_smem = 48 * 1024;
  if (smem > attr_smem) {
    SGA_CUDA(cudaFuncSetAttribute(sga::selftest_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  sga::selftest_umma_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, Ncols, K, kind);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}
