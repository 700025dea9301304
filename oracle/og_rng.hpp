// og_rng.hpp -- TEST INFRASTRUCTURE (oracle).  Not part of the product.
//
// The reference draws from OCaml's global `Random` (111 call sites, e.g.
// mcmc.ml:47, interpolate_pdf.ml:115); its stream is version dependent and is
// NOT reproduced.  The oracle and the GPU path instead share one
// *specification* of a counter-based stream, implemented twice (here in plain
// C++, in mcmc_ocaml_b200/csrc/rng.cuh for the device), so that stochastic
// entry points can be compared draw-for-draw on the same seed:
//
//   call key   (k0,k1) = Philox4x32-10(ctr = (epoch_lo, epoch_hi, 'mcmc', 0),
//                                      key = (seed_lo, seed_hi)).words[0..1]
//   block      B(purpose, g, step, blk) = Philox4x32-10(
//                  ctr = (blk, step_lo, g_lo,
//                         g_hi16 | purpose << 16 | step_hi8 << 24), key = call key)
//   draw j of (purpose, g, step) = the 64-bit lane  (w[2(j&1)] << 32 | w[2(j&1)+1])
//                  of block j >> 1
//   Random.float 1.0  ->  u52 = bits(0x3FF<<52 | lane & (2^52-1)) - 1.0   in [0,1)
//                         (mantissa = low 20 bits of the first word : second word)
//   Random.int n      ->  mulhi64(lane, n)
//
// Philox4x32-10: Salmon et al., "Parallel random numbers: as easy as 1, 2, 3"
// (SC'11); checked against the Random123 known-answer vectors in
// tests/test_oracle_rng.py.
#pragma once
#include <cstdint>
#include <cstring>

namespace og {

struct Philox {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  static constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  static inline void block(const uint32_t ctr[4], const uint32_t key[2],
                           uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
      uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
      uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0;
      uint32_t h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
      uint32_t n0 = h1 ^ c1 ^ k0, n1 = l1, n2 = h0 ^ c3 ^ k1, n3 = l0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
};

enum Purpose : uint32_t {
  P_MH = 1, P_RJ = 2, P_RJ_INIT = 3, P_DRAW = 4, P_NEST_INIT = 5,
  P_NEST_MCMC = 6, P_NEST_START = 7, P_POST = 8, P_BOOT = 9
};

struct CallKey { uint32_t k[2]; };

inline CallKey derive_key(uint64_t seed, uint64_t epoch) {
  uint32_t ctr[4] = {(uint32_t)epoch, (uint32_t)(epoch >> 32), 0x6d636d63u, 0u};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t out[4];
  Philox::block(ctr, key, out);
  return CallKey{{out[0], out[1]}};
}

// Sequential cursor over the draws of one (purpose, g, step).
struct Rng {
  uint32_t key[2];
  uint32_t ctr[4];
  uint32_t w[4];
  uint32_t j = 0;  // next draw index
  Rng(const CallKey &ck, uint32_t purpose, uint64_t g, uint64_t step) {
    key[0] = ck.k[0]; key[1] = ck.k[1];
    ctr[0] = 0;
    ctr[1] = (uint32_t)step;
    ctr[2] = (uint32_t)g;
    ctr[3] = (uint32_t)((g >> 32) & 0xFFFFu) | ((purpose & 0xFFu) << 16) |
             (uint32_t)(((step >> 32) & 0xFFu) << 24);
  }
  inline uint64_t lane() {
    if ((j & 1u) == 0u) { ctr[0] = j >> 1; Philox::block(ctr, key, w); }
    uint32_t a = w[2 * (j & 1u)], b = w[2 * (j & 1u) + 1];
    ++j;
    return ((uint64_t)a << 32) | b;
  }
  // Random.float 1.0
  inline double uniform() {
    uint64_t bits = (0x3FFull << 52) | (lane() & 0xFFFFFFFFFFFFFull);
    double d; std::memcpy(&d, &bits, 8);
    return d - 1.0;
  }
  // 1 + Random.float 1.0, in [1, 2): the raw mantissa pattern, no subtraction
  inline double uniform12() {
    uint64_t bits = (0x3FFull << 52) | (lane() & 0xFFFFFFFFFFFFFull);
    double d; std::memcpy(&d, &bits, 8);
    return d;
  }
  // Random.int n
  inline uint64_t below(uint64_t n) {
    return (uint64_t)(((unsigned __int128)lane() * n) >> 64);
  }
};

}  // namespace og
