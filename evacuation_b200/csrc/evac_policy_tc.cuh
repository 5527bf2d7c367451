// Layer 1 of the policy heads on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
//   H1[E, 128] = tanh( X[E, K] * W1[K, 128] + b1 )        X = transformer embedding (K = S * D = 372 at the reference shape),
//                                                         W1 = [critic.0.weight^T | actor_mean.0.weight^T]
//   [src/agents/networks/rpo_linear_agent_network.py:23-42 -- the one dense contraction on the rollout path]
//
// One CTA = 128 environments x NT columns (NT = 64: two CTAs per 128 environments, for batches that would otherwise leave
// SMs idle; NT = 128 for large batches); K in chunks of 32 (one 128-byte swizzle row of float32).  float32 fidelity from
// kind::tf32 MMAs (10-bit mantissa operands) by the 3xTF32 split:  x = xh + xl, w = wh + wl (xh / wh = the value with the
// 13 low mantissa bits cleared, xl / wl = the float32 remainder, truncated by the tensor core to its top 10 bits):
//   x w  ~=  xh wh + xl wh + xh wl          (error ~2^-21 per product; the xl wl term is below float32 resolution)
// accumulated in float32 in tensor memory.  Per chunk and CTA: 4 k-steps (UMMA_K = 8) x 3 products = 12 tcgen05.mma
// (M = 128, N = NT) issued by ONE thread of a dedicated warp; operands in shared memory in the canonical K-major
// SWIZZLE_128B layout (8-row x 128-byte atoms, 16-byte chunk index XOR row-in-atom) -- X is split and laid out by the four
// producer warps (global loads of chunk c + 1 in flight while chunk c is stored), the weight tiles are split and swizzled
// once on the host (evac_policy_load_weights) and arrive as ONE bulk copy per chunk (cp.async.bulk + mbarrier
// complete_tx).  Ring of TC_STAGES stages: full[s] (128 producer arrivals + the bulk copy's bytes) / empty[s]
// (tcgen05.commit).  Epilogue: tcgen05.ld 32x32b (warp w owns TMEM lanes 32w .. 32w + 31 = environments), + bias, tanh,
// float4 stores of H1; evac_policy_heads_kernel continues from H1 (HArgs::h1_in).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace evacp {

constexpr int TC_M = 128, TC_KC = 32;
constexpr int TC_COLS = 128;                                 // = HD_COLS
constexpr int TC_THREADS = 160;                              // warps 0-3: producers + epilogue, warp 4: MMA issue + TMEM allocation
constexpr int TC_XTILE_BYTES = TC_M * TC_KC * 4;             // 16 KB: one X tile (128 rows x 128 bytes)
template <int NT> struct TCShape {
  static constexpr int WTILE_BYTES = NT * TC_KC * 4;         // one W tile (NT rows x 128 bytes)
  static constexpr int STAGE_BYTES = 2 * TC_XTILE_BYTES + 2 * WTILE_BYTES;   // X hi | X lo | W hi | W lo
  static constexpr int STAGES = NT == 64 ? 4 : 3;
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
};

struct TCArgs {
  int E, K, chunks;        // chunks = ceil(K / 32)
  const float* emb;        // [E, K] float32, rows 16-byte aligned (K % 4 == 0)
  const float* w1tc;       // [chunks][128 / NT][hi tile | lo tile], each tile NT (column) rows x 32 k, swizzled
  const float* b1;         // [128]
  float* h1;               // [E, 128] out
};

// byte offset of element (row r, 16-byte chunk c4) inside a K-major SWIZZLE_128B tile
__host__ __device__ __forceinline__ uint32_t tc_swizzle(int r, int c4) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c4 ^ (r & 7)) << 4)); }
// the value with its 13 low mantissa bits cleared: exactly representable as TF32
__host__ __device__ __forceinline__ float tc_hi(float x) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
#else
  uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y;
#endif
}

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(tc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(tc_smem_u32(bar)), "r"(parity) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in [0,14), leading
// byte offset (unused for swizzled K-major layouts; 1) in [16,30), stride byte offset = 1024 B (next 8-row atom) >> 4 in
// [32,46), descriptor version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64)
__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor), kind::tf32: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major,
// N >> 3 in [17,23), M >> 4 in [24,29)
template <int NT>
__host__ __device__ constexpr uint32_t tc_idesc() { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24); }

__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {  // implies tcgen05.fence::before_thread_sync
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}

template <int NT>
__global__ void __launch_bounds__(TC_THREADS, 1) evac_policy_l1_tc_kernel(const __grid_constant__ TCArgs a) {
  using SH = TCShape<NT>;
  constexpr int STAGES = SH::STAGES;
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);  // swizzle atoms: 1024-byte aligned
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * SH::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* accum = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int e0 = blockIdx.x * TC_M;
  const int ny = TC_COLS / NT, y = blockIdx.y;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { tc_mbar_init(&full[s], 128); tc_mbar_init(&empty[s], 1); }
    tc_mbar_init(accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {  // NT TMEM columns of float32 accumulators (power of two >= 32), one warp allocates and later frees
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "n"(NT) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    // ---- MMA issue: one thread, 12 tcgen05.mma per chunk, stage handed back through tcgen05.commit
    if (lane == 0) {
      constexpr uint32_t IDESC = tc_idesc<NT>();
      for (int c = 0; c < a.chunks; ++c) {
        const int s = c % STAGES;
        tc_mbar_wait(&full[s], (uint32_t)((c / STAGES) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t base = tc_smem_u32(smem + s * SH::STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < TC_KC / 8; ++k) {  // UMMA_K = 8 float32 = 32 bytes along the swizzled row
          const uint64_t xh = tc_desc(base + k * 32), xl = tc_desc(base + TC_XTILE_BYTES + k * 32);
          const uint64_t wh = tc_desc(base + 2 * TC_XTILE_BYTES + k * 32), wl = tc_desc(base + 2 * TC_XTILE_BYTES + SH::WTILE_BYTES + k * 32);
          tc_mma(tmem, xh, wh, IDESC, (c > 0 || k > 0) ? 1u : 0u);
          tc_mma(tmem, xl, wh, IDESC, 1u);
          tc_mma(tmem, xh, wl, IDESC, 1u);
        }
        tc_commit(&empty[s]);
      }
      tc_commit(accum);
    }
  } else {
    // ---- producers: split X into hi / lo tiles (registers hold chunk c + 1 while chunk c is stored), request the weight tiles
    const int kq = a.K >> 2;  // float4 per row of X
    auto load_x = [&](int c, float4 (&x)[8]) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = i * 128 + tid, r = idx >> 3, c4 = idx & 7;
        const int e = e0 + r, q = c * 8 + c4;
        x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < a.E && q < kq) x[i] = __ldg(reinterpret_cast<const float4*>(a.emb + (size_t)e * a.K + 4 * q));
      }
    };
    float4 xr[8], xn[8];
    load_x(0, xr);
    for (int c = 0; c < a.chunks; ++c) {
      const int s = c % STAGES, use = c / STAGES;
      if (c + 1 < a.chunks) load_x(c + 1, xn);
      if (use > 0) tc_mbar_wait(&empty[s], (uint32_t)((use - 1) & 1));  // the MMAs that read this stage have completed
      uint8_t* st = smem + s * SH::STAGE_BYTES;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = i * 128 + tid, r = idx >> 3, c4 = idx & 7;
        const float4 x = xr[i];
        float4 hi, lo;
        hi.x = tc_hi(x.x); lo.x = x.x - hi.x;
        hi.y = tc_hi(x.y); lo.y = x.y - hi.y;
        hi.z = tc_hi(x.z); lo.z = x.z - hi.z;
        hi.w = tc_hi(x.w); lo.w = x.w - hi.w;
        const uint32_t off = tc_swizzle(r, c4);
        *reinterpret_cast<float4*>(st + off) = hi;
        *reinterpret_cast<float4*>(st + TC_XTILE_BYTES + off) = lo;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core (async proxy)
      if (tid == 0) {
        tc_mbar_arrive_expect_tx(&full[s], 2 * SH::WTILE_BYTES);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc_smem_u32(st + 2 * TC_XTILE_BYTES)),
                     "l"(a.w1tc + ((size_t)c * ny + y) * (2 * SH::WTILE_BYTES / 4)), "r"(2 * SH::WTILE_BYTES), "r"(tc_smem_u32(&full[s]))
                     : "memory");
      } else {
        tc_mbar_arrive(&full[s]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) xr[i] = xn[i];
    }
    // ---- epilogue: TMEM lane = environment, column = hidden unit
    tc_mbar_wait(accum, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int e = e0 + tid;
#pragma unroll 1
    for (int j = 0; j < NT / 32; ++j) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(j * 32);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, "
          "%29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
            "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
            "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
            "=r"(v[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (e < a.E) {
        const int col0 = y * NT + j * 32;
        float* dst = a.h1 + (size_t)e * TC_COLS + col0;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(a.b1 + col0 + i));
          *reinterpret_cast<float4*>(dst + i) = make_float4(tanhf(__uint_as_float(v[i]) + b.x), tanhf(__uint_as_float(v[i + 1]) + b.y),
                                                            tanhf(__uint_as_float(v[i + 2]) + b.z), tanhf(__uint_as_float(v[i + 3]) + b.w));
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(NT) : "memory");
  }
}

}  // namespace evacp
