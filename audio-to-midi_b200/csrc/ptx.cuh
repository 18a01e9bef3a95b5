// Thin inline-PTX wrappers for the sm_100a features the kernels use: mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and the UMMA descriptors.
// Nothing here is portable: compile with -gencode arch=compute_100a,code=sm_100a only.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace a2m {

// Optional in-kernel timelines (build with -DA2M_FFN_TIMING): CTA 0 of the instrumented kernels records clock64() at its
// phase boundaries, read back with a2m_debug_read_timing() (tools/ffn_timeline.py).  Compiled out of the product build.
#ifdef A2M_FFN_TIMING
__device__ long long g_ffn_timing[128];
#endif

// ---------------------------------------------------------------- tensor-core operand format (build variant)
// tcgen05 kind::f16 takes bf16 OR IEEE binary16 operands at the same rate.  The library is built twice from the same sources:
//   libaudio2midi_b200.so      bf16 operands (8-bit significand, fp32 exponent range): training, and inference when asked;
//   libaudio2midi_b200_f16.so  -DA2M_OP_F16: binary16 operands (11-bit significand, 8x smaller rounding) for INFERENCE -- the
//                              reference infers in fp32 and trains in fp16 (infer.py:234, train.py:36-37), so fp16 range is known
//                              to suffice for its checkpoints; the residual stream, accumulators and LN / softmax statistics are
//                              fp32 in both.  a2m_train_init is refused in this variant (the backward's unpack helpers are bf16).
// Only three things differ: how two floats are rounded into a 16-bit pair (op2_rn / op1_rn), the operand-format bits of the UMMA
// instruction descriptor (kOpFmt), and the host-side weight image (f32_to_op16 in a2m_api.cu).  `__nv_bfloat16*` pointers are
// 16-bit carriers in both variants.
#ifdef A2M_OP_F16
constexpr uint32_t kOpFmt = 0;   // instruction-descriptor A/B format: 0 = f16
__device__ __forceinline__ __nv_bfloat162 op2_rn(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<__nv_bfloat162*>(&h);
}
__device__ __forceinline__ __nv_bfloat16 op1_rn(float x) {
  __half h = __float2half_rn(x);
  return *reinterpret_cast<__nv_bfloat16*>(&h);
}
__device__ __forceinline__ float op2_sum(__nv_bfloat162 v) {   // lo + hi of a rounded pair, as floats
  const float2 f = __half22float2(*reinterpret_cast<__half2*>(&v));
  return f.x + f.y;
}
#else
constexpr uint32_t kOpFmt = 1;   // 1 = bf16
__device__ __forceinline__ __nv_bfloat162 op2_rn(float lo, float hi) { return __floats2bfloat162_rn(lo, hi); }
__device__ __forceinline__ __nv_bfloat16 op1_rn(float x) { return __float2bfloat16_rn(x); }
__device__ __forceinline__ float op2_sum(__nv_bfloat162 v) { return __low2float(v) + __high2float(v); }
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Programmatic dependent launch: let the next kernel in the stream start its prologue early / wait until the
// previous kernel's writes are visible.  Both are no-ops when the launch did not opt in.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// Copies NVEC 16-byte vectors of CONSTANT data (weights, parameter images) from global to shared memory with every load of
// a thread issued before its first store: a plain `for (i = tid; i < n; i += threads) s[i] = __ldg(g + i)` loop compiles to
// one dependent global round trip per iteration, which showed up as 10-25 % of the stall samples of the short kernels.
template <int NVEC, int THREADS>
__device__ __forceinline__ void copy_const_to_smem(void* smem_dst, const void* gmem_src, int tid) {
  constexpr int PER = (NVEC + THREADS - 1) / THREADS;
  uint4 v[PER];
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int i = tid + k * THREADS;
    if (i < NVEC) v[k] = __ldg(reinterpret_cast<const uint4*>(gmem_src) + i);
  }
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int i = tid + k * THREADS;
    if (i < NVEC) reinterpret_cast<uint4*>(smem_dst)[i] = v[k];
  }
}

// ---------------------------------------------------------------- dropout (training)
// eqx.nn.Dropout (model.py:224, 335): keep with probability 1 - p, scale kept values by 1 / (1 - p).  Counter based:
// the decision for element `idx` of dropout site `site` is a pure function of (seed, site, idx), so the backward
// kernels regenerate the forward's mask instead of storing it.  The parameters live in device memory (one struct per
// handle, rewritten before every forward) so that CUDA graphs need no re-capture when the seed changes.
struct DropParams {
  uint32_t seed;
  uint32_t thresh;    // drop iff hash < thresh;  thresh = round(p * 2^32), 0 = dropout off
  float inv_keep;     // 1 / (1 - p)
  uint32_t pad;
};
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {   // "lowbias32" integer finaliser
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t drop_key(uint32_t seed, uint32_t site) { return mix32(seed ^ (site * 0x85ebca6bu + 0x9e3779b9u)); }
// multiplier of element idx: 1 / (1 - p) if kept, 0 if dropped
__device__ __forceinline__ float drop_mul(uint32_t key, uint32_t idx, uint32_t thresh, float inv_keep) {
  return mix32((idx * 0x9e3779b1u) ^ key) >= thresh ? inv_keep : 0.f;
}
enum DropSite : uint32_t { DROP_FFN = 0, DROP_GLOBAL = 1, DROP_LOCAL = 2 };   // site id = layer * 4 + kind

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  // generic-proxy writes to shared memory -> visible to the async proxy (TMA / tcgen05.mma)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must surface as a trapped kernel (cudaErrorLaunchFailure), never as
// a hung GPU.  ~4e9 SM cycles is about two seconds; no legitimate wait in this library is near that.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled load, coordinates {c0 = innermost (element), c1 = row}; completes on `bar`.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// 3-D tiled load {c0, c1, c2}.
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// 2-D tiled store smem -> global (bulk async-group completion); rows/cols outside the tensor are clipped.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// waits until at most N of the most recent bulk groups of this thread are still READING shared memory
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ---------------------------------------------------------------- tcgen05
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(kCols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // whole warp (the allocating one)
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(kCols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T ; bf16 x bf16 -> fp32.  One thread issues for the whole CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives when all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// TMEM -> registers: warp w may touch lanes [32*(w%4), 32*(w%4)+32); thread t gets lane base+t,
// register j = column (addr.col + j).
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor for a K-major bf16 tile stored as rows of 128 bytes (64 bf16)
// with the 128-byte swizzle (what a TMA load with CU_TENSOR_MAP_SWIZZLE_128B produces, and what
// sw128_offset() below produces for tiles written by threads).  The tile base must be 1024 B aligned.
//   bits [0,14)  start address >> 4
//   bits [16,30) leading byte offset >> 4   (unused for swizzled K-major; 1 by convention)
//   bits [32,46) stride byte offset >> 4    (1024 B between 8-row groups)
//   bits [46,48) descriptor version = 1 (Blackwell)
//   bits [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// Advancing K by 16 bf16 (= 32 bytes) inside the 128-byte swizzle atom is a plain add on the
// start-address field.
__device__ __forceinline__ uint64_t umma_desc_advance_k(uint64_t d, uint32_t bytes) { return d + (bytes >> 4); }

// Instruction descriptor: kind::f16, A = B = bf16 (or f16 in the A2M_OP_F16 build), D = fp32, both operands K-major, dense.
//   bits [4,6) D format (1 = f32); [7,10) A format (0 = f16, 1 = bf16); [10,13) B format (likewise);
//   bit 15/16 A/B major (0 = K); [17,23) N >> 3; [24,29) M >> 4.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t m, uint32_t n) {
  return (1u << 4) | (kOpFmt << 7) | (kOpFmt << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// Same, with the B operand MN-major: B is stored as [K rows][N contiguous] (e.g. V[key][d] for P.V), i.e. the
// transpose bit of B (bit 16) is set.  With 64 bf16 (128 B) of N per row and the 128-byte swizzle, the smem tile
// uses the same descriptor as a K-major tile; a K step of 16 advances the start address by 16 rows = 2048 B.
__host__ __device__ constexpr uint32_t umma_idesc_bf16_bmn(uint32_t m, uint32_t n) {
  return umma_idesc_bf16(m, n) | (1u << 16);
}

// Byte offset of element (row, col_bf16) inside a [rows x 64] bf16 tile in the canonical
// 128-byte-swizzled K-major layout (Swizzle<3,4,3>): the 16-byte chunk index is XORed with row % 8.
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t col) {
  uint32_t chunk = (col >> 3) ^ (row & 7u);
  return row * 128u + chunk * 16u + (col & 7u) * 2u;
}

}  // namespace a2m
