// Counter-based random draws for the batched ball environment (sm_100a).
//
// The reference draws from the process-global np.random MT19937 stream in textual order
// (gym_ballenv/envs/ballenv_env.py:24-25,115-118,332,340,345,352).  Here every draw has an
// address and its 32-bit word is Philox4x32-10(counter = address, key = seed) or, in parity
// mode, a word injected through ballenv_set_draw_tape().  randint(n) = mulhi(word, n).
//
//   counter = { global env id, tick | episode, block, stream }
//   step  draws : stream 1, c1 = tick,    block = j >> 2, word j & 3  (one word per moving obstacle;
//                 the second draw of the same obstacle uses the low half of word * 100)
//   reset draws : stream 2, c1 = episode, block = kind << 28 | item/attempt bits
//
// CPU restatement used by the tests: oracle/draws.py.
#pragma once
#include <stdint.h>

namespace ballenv {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

constexpr uint32_t kStreamStep = 1;
constexpr uint32_t kStreamReset = 2;
constexpr uint32_t kStreamAction = 3;

constexpr uint32_t kResetHead = 0;
constexpr uint32_t kResetStatic = 1;
constexpr uint32_t kResetDynamic = 2;
constexpr uint32_t kResetAgentRedraw = 3;

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(kPhiloxM0, c0), lo0 = kPhiloxM0 * c0;
    const uint32_t hi1 = __umulhi(kPhiloxM1, c2), lo1 = kPhiloxM1 * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += kPhiloxW0;
    k1 += kPhiloxW1;
  }
  return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ uint32_t pick_word(const uint4& b, uint32_t i) {
  return i == 0 ? b.x : (i == 1 ? b.y : (i == 2 ? b.z : b.w));
}

// 53-bit double in [0, 1) from two words, as NumPy's legacy random_sample does.
__device__ __forceinline__ double ranf_from_words(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

}  // namespace ballenv
