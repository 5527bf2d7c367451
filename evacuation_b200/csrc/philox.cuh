// Philox4x32-10 counter-based generator (Salmon et al., "Parallel random numbers: as easy as
// 1, 2, 3", SC'11).  Replaces the reference's global MT19937 stream (area.py:124,
// pedestrians.py:17-18): every (seed, stream, env, episode, step, pedestrian) tuple owns its
// own random words, so results do not depend on how environments are sharded over CTAs/GPUs.
// tests/evac_testlib.py holds the NumPy restatement the parity tests compare against.
#pragma once
#include <stdint.h>

namespace evac {

enum : uint32_t { STREAM_NOISE = 0, STREAM_RESET = 1, STREAM_AGENT = 2 };

struct Philox4 {
  uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
  const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
  c[0] = hi1 ^ c[1] ^ k0;
  c[1] = lo1;
  c[2] = hi0 ^ c[3] ^ k1;
  c[3] = lo0;
}

// counter = (c0, c1, c2, c3), key = (k0, k1)
__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
  uint32_t c[4] = {c0, c1, c2, c3};
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return Philox4{c[0], c[1], c[2], c[3]};
}

// Stream layout used everywhere in this library:
//   4x32 (reset layout): counter = (pedestrian, 0, episode index, global env index),
//                        key = (seed_lo ^ stream * 0x9E3779B9, seed_hi)
//   2x32 (noise, agent): see evac_noise_block / evac_agent_block below
__host__ __device__ __forceinline__ Philox4 evac_random(uint64_t seed, uint32_t stream, uint32_t env, uint32_t episode,
                                                        uint32_t now, uint32_t ped) {
  return philox4x32_10(ped, now, episode, env, (uint32_t)seed ^ (stream * 0x9E3779B9u), (uint32_t)(seed >> 32));
}

// Philox2x32-10: one 32x32 multiply per round -- the per-step angular noise (one block per lane = the lane's two
// pedestrians) and the RandomAgent action use it; the (rare) reset layout keeps the 4x32 generator above.
__host__ __device__ __forceinline__ uint2 philox2x32_10(uint32_t c0, uint32_t c1, uint32_t k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p = (uint64_t)0xD256D193u * c0;
    const uint32_t hi = (uint32_t)(p >> 32), lo = (uint32_t)p;
    c0 = hi ^ k ^ c1;
    c1 = lo;
    k += 0x9E3779B9u;
  }
  return make_uint2(c0, c1);
}

// 32-bit key of the 2x32 streams: for a fixed seed it is a bijection of the episode index (odd multiplier), so
// (env, episode, step, block) -> (counter, key) never collides within one run.
__host__ __device__ __forceinline__ uint32_t evac_key32(uint64_t seed, uint32_t stream, uint32_t episode) {
  return (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x9E3779B9u) ^ (episode * 0xBB67AE85u) ^ (stream * 0x85EBCA6Bu);
}

// Angular noise [area.py:124]: pedestrian i reads word (i >> 5) & 1 of block (i & 31) | (i >> 6) << 5, i.e. pedestrians i and
// i + 32 of every group of 64 share a block.  counter = ((block & 2047) | now << 11, global env): the low 11 block bits sit
// next to the 21-bit step index; the upper block bits (block >= 2048 <=> pedestrian >= 4096, crowds up to 32 768) go into
// the KEY through an odd multiplier -- for a fixed (env, episode, step) every block still owns its own (counter, key).
// (They must not spill into the step field: block | now << 11 would alias pedestrian i + 4096 at step t with pedestrian i
// at step t + 1.)
__host__ __device__ __forceinline__ uint2 evac_noise_block(uint64_t seed, uint32_t env, uint32_t episode, uint32_t now, uint32_t block) {
  return philox2x32_10((block & 2047u) | (now << 11), env, evac_key32(seed, STREAM_NOISE, episode) ^ ((block >> 11) * 0xC2B2AE35u));
}
__host__ __device__ __forceinline__ uint32_t evac_noise_block_of(uint32_t i) { return (i & 31u) | ((i >> 6) << 5); }
__host__ __device__ __forceinline__ uint32_t evac_noise_word_of(uint32_t i) { return (i >> 5) & 1u; }

// RandomAgent action (action_space.sample(), random_agent.py:8-9): words (x, y) of one block per (env, episode, step).
__host__ __device__ __forceinline__ uint2 evac_agent_block(uint64_t seed, uint32_t env, uint32_t episode, uint32_t now) {
  return philox2x32_10(now << 11, env, evac_key32(seed, STREAM_AGENT, episode));
}

// 24-bit uniform in [0,1): exactly representable in float32, so NumPy reproduces it bit for bit.
__host__ __device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * 5.9604644775390625e-08f; }

}  // namespace evac
