// Host-side helper: advance torch's CPU generator (mt19937, ATen/core/MT19937RNGEngine.h) by n 32-bit draws WITHOUT
// producing them.  The reference draws its jitter sequentially over the image from the global CPU generator
// (implicit_surface.py:276,305,174; quirk Q1), so a rank that renders a ray shard bit-identically has to consume the
// draws of every other rank's chunks: 7.4 M draws per 576x800 image.  torch.rand() of that many floats costs ~20 ms
// (tempering, int -> float, the store); the state transition alone is ~12 k "twists" of 624 words that vectorise.
// Operates on the fields of the state blob torch.get_rng_state() returns (CPUGeneratorImplStateLegacy: left, next,
// state[624] as 64-bit words).  One float32 of torch.rand consumes one draw.
#include <stdint.h>
#include <string.h>

#include "surf_internal.cuh"

#define MT_N 624
#define MT_M 397

static inline uint32_t mt_twist(uint32_t u, uint32_t v) {
  return (((u & 0x80000000u) | (v & 0x7fffffffu)) >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
}
// ATen's mt19937::next_state(): p[i] = p[i + M] ^ twist(p[i], p[i + 1]), indices mod N
static void mt_next_state(uint32_t* __restrict__ s) {
  for (int i = 0; i < MT_N - MT_M; ++i) s[i] = s[i + MT_M] ^ mt_twist(s[i], s[i + 1]);
  for (int i = MT_N - MT_M; i < MT_N - 1; ++i) s[i] = s[i + MT_M - MT_N] ^ mt_twist(s[i], s[i + 1]);
  s[MT_N - 1] = s[MT_M - 1] ^ mt_twist(s[MT_N - 1], s[0]);
}

extern "C" int surf_mt19937_skip(uint64_t* h_state624, int32_t* h_left, uint64_t* h_next, uint64_t n_draws) {
  SURF_CHECK_ARG(h_state624 && h_left && h_next, "null pointer");
  int64_t left = *h_left;
  uint64_t next = *h_next;
  SURF_CHECK_ARG(left >= 1 && left <= MT_N && next <= MT_N, "not an mt19937 state");
  // a draw does: if (--left == 0) next_state() [left = 624, next = 0]; y = state[next++]
  const uint64_t avail = (uint64_t)(left - 1);
  if (n_draws <= avail) {
    *h_left = (int32_t)(left - (int64_t)n_draws);
    *h_next = next + n_draws;
    return 0;
  }
  uint64_t n = n_draws - avail;                  // draws that need fresh state
  uint32_t s[MT_N];
  for (int i = 0; i < MT_N; ++i) s[i] = (uint32_t)h_state624[i];
  const uint64_t twists = (n + MT_N - 1) / MT_N;
  for (uint64_t t = 0; t < twists; ++t) mt_next_state(s);
  const uint64_t used = n - (twists - 1) * MT_N;  // 1 .. 624 draws taken from the last state
  for (int i = 0; i < MT_N; ++i) h_state624[i] = s[i];
  *h_next = used;
  *h_left = (int32_t)(MT_N + 1 - used);
  return 0;
}
