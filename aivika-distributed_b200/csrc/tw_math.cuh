// tw_math.cuh -- bit-exact scalar arithmetic of the engine (device + host).
//
// Every binary64 operation is issued as one correctly rounded IEEE operation
// (__dmul_rn / __dadd_rn / __ddiv_rn on the device, plain operators on the host where
// the translation unit is compiled with -ffp-contract=off), so that timestamps produced
// on the GPU are bit-identical to the CPU oracle's (oracle/oracle_math.h is an
// independent restatement of the same definitions; tests/ compare the two).
//
// Reference semantics followed:
//   exponential variate  `- log x * mu`   tests/MachRep2DistributedReproducible.hs:81-86,
//                                         Simulation/Aivika/Distributed/Optimistic/Generator.hs:52
//   StdGen / random      package `random` 1.1 (mkStdGen, next, random :: Double), call sites
//                        tests/MachRep2DistributedReproducible.hs:78,84
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define TW_HD __host__ __device__ __forceinline__
#else
#define TW_HD inline
#endif

namespace twb {

#if defined(__CUDA_ARCH__)
TW_HD double fmul(double a, double b) { return __dmul_rn(a, b); }
TW_HD double fadd(double a, double b) { return __dadd_rn(a, b); }
TW_HD double fsub(double a, double b) { return __dsub_rn(a, b); }
TW_HD double fdiv(double a, double b) { return __ddiv_rn(a, b); }
TW_HD uint64_t d2bits(double x) { return (uint64_t)__double_as_longlong(x); }
TW_HD double bits2d(uint64_t u) { return __longlong_as_double((long long)u); }
TW_HD uint64_t mulhi64(uint64_t a, uint64_t b) { return __umul64hi(a, b); }
#else
TW_HD double fmul(double a, double b) { return a * b; }
TW_HD double fadd(double a, double b) { return a + b; }
TW_HD double fsub(double a, double b) { return a - b; }
TW_HD double fdiv(double a, double b) { return a / b; }
TW_HD uint64_t d2bits(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
TW_HD double bits2d(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
TW_HD uint64_t mulhi64(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }
#endif

constexpr uint64_t kGolden = 0x9E3779B97F4A7C15ull;
constexpr uint64_t kMulA   = 0xBF58476D1CE4E5B9ull;
constexpr uint64_t kMulB   = 0x94D049BB133111EBull;
constexpr uint64_t kCtrMul = 0xD1B54A32D192ED03ull;

// splitmix64 output function
TW_HD uint64_t mix_fin(uint64_t z)
{
    z = (z ^ (z >> 30)) * kMulA;
    z = (z ^ (z >> 27)) * kMulB;
    return z ^ (z >> 31);
}
// inverse of mix_fin (xor-shifts and odd multipliers are bijections of 64-bit words)
constexpr uint64_t kMulAInv = 0x96DE1B173F119089ull;
constexpr uint64_t kMulBInv = 0x319642B2D24D8EC3ull;
TW_HD uint64_t mix_fin_inv(uint64_t z)
{
    z ^= z >> 31; z ^= z >> 62;
    z *= kMulBInv;
    z ^= z >> 27; z ^= z >> 54;
    z *= kMulAInv;
    z ^= z >> 30; z ^= z >> 60;
    return z;
}
TW_HD uint64_t lp_stream_key(uint64_t seed, uint32_t lp) { return mix_fin(seed ^ (((uint64_t)lp + 1ull) * kGolden)); }
TW_HD uint64_t stream_draw(uint64_t key, uint64_t ctr) { return mix_fin(key + ctr * kCtrMul); }

// uniform on (0,1): 52 bits + 1/2, exact
TW_HD double uniform01(uint64_t d) { return fmul(fadd((double)(d >> 12), 0.5), 0x1p-52); }

// natural log of a positive normal double (0 -> -inf); fixed operation order, no FMA.
TW_HD double log_exact(double x)
{
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double c1 = 6.666666666666735130e-01, c2 = 3.999999999940941908e-01, c3 = 2.857142874366239149e-01,
                 c4 = 2.222219843214978396e-01, c5 = 1.818357216161805012e-01, c6 = 1.531383769920937332e-01,
                 c7 = 1.479819860511658591e-01;
    uint64_t b = d2bits(x);
    if (b == 0) return bits2d(0xFFF0000000000000ull);
    int64_t k = (int64_t)(b >> 52) - 1023;
    uint64_t m = b & 0x000FFFFFFFFFFFFFull;
    if (m > 0x6A09E667F3BCCull) { k += 1; b = m | 0x3FE0000000000000ull; }
    else                        {         b = m | 0x3FF0000000000000ull; }
    double f = fsub(bits2d(b), 1.0);
    double dk = (double)k;
    double s = fdiv(f, fadd(2.0, f));
    double z = fmul(s, s);
    double w = fmul(z, z);
    double t1 = fmul(w, fadd(c2, fmul(w, fadd(c4, fmul(w, c6)))));
    double t2 = fmul(z, fadd(c1, fmul(w, fadd(c3, fmul(w, fadd(c5, fmul(w, c7)))))));
    double R = fadd(t2, t1);
    double hfsq = fmul(fmul(0.5, f), f);
    double inner = fadd(fmul(s, fadd(hfsq, R)), fmul(dk, ln2_lo));
    return fsub(fmul(dk, ln2_hi), fsub(fsub(hfsq, inner), f));
}

TW_HD double expo_variate(double u, double mu) { return -fmul(log_exact(u), mu); }

// order-dependent committed-trace hash
TW_HD uint64_t trace_fold(uint64_t h, double recv_ts, uint32_t src, uint64_t ctr)
{
    uint64_t x = d2bits(recv_ts) ^ (ctr * kMulB) ^ ((uint64_t)src * kMulA);
    return mix_fin((h ^ x) + kGolden);
}

// the fold is invertible: reverse computation restores the hash before an undone event
TW_HD uint64_t trace_unfold(uint64_t h, double recv_ts, uint32_t src, uint64_t ctr)
{
    uint64_t x = d2bits(recv_ts) ^ (ctr * kMulB) ^ ((uint64_t)src * kMulA);
    return (mix_fin_inv(h) - kGolden) ^ x;
}

// ---- StdGen (random-1.1: L'Ecuyer combined MLCG) ----
struct StdGen { int32_t s1, s2; };

TW_HD StdGen mk_std_gen(int64_t seed)
{
    int32_t s = (int32_t)((uint32_t)seed & 0x7FFFFFFFu);
    int32_t q = s / 2147483562, r = s % 2147483562;
    StdGen g; g.s1 = r + 1; g.s2 = (q % 2147483398) + 1;
    return g;
}
TW_HD int32_t std_next(StdGen &g)
{
    int32_t k = g.s1 / 53668;
    int32_t a = 40014 * (g.s1 - k * 53668) - k * 12211;
    if (a < 0) a += 2147483563;
    int32_t k2 = g.s2 / 52774;
    int32_t b = 40692 * (g.s2 - k2 * 52774) - k2 * 3791;
    if (b < 0) b += 2147483399;
    g.s1 = a; g.s2 = b;
    int32_t z = a - b;
    return z < 1 ? z + 2147483562 : z;
}
// random :: StdGen -> (Double, StdGen): three `next` draws in base 2147483562, low 53 bits / 2^53
TW_HD double std_random_double(StdGen &g)
{
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) v = v * 2147483562ull + ((uint64_t)std_next(g) - 1ull);
    return fmul((double)(v & ((1ull << 53) - 1ull)), 0x1p-53);
}

}  // namespace twb
