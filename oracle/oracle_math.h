/*
 * oracle_math.h -- TEST INFRASTRUCTURE (CPU oracle).  Not part of the product path.
 *
 * Bit-exact scalar arithmetic used by the CPU oracle: the counter-based draw
 * function, the uniform (0,1) mapping, a software natural logarithm and the
 * StdGen generator of the `random` package (1.1 regime).
 *
 * The reference (aivika-distributed) takes its arithmetic from un-vendored
 * third-party packages:
 *   - exponential variates: `generateExponential01 g mu = (\x -> - log x * mu)`
 *     (aivika-transformers; call site Simulation/Aivika/Distributed/Optimistic/Generator.hs:52)
 *     and `randomRefExponential` (tests/MachRep2DistributedReproducible.hs:81-86);
 *   - `mkStdGen`, `random :: StdGen -> (Double, StdGen)` (package `random`,
 *     call sites tests/MachRep2DistributedReproducible.hs:78,84).  The cabal bound
 *     `random >= 1.0.0.3` admits both the L'Ecuyer (<= 1.1) and SplitMix (>= 1.2)
 *     generators; this oracle restates the published 1.1 algorithm.  PARITY UNPINNED:
 *     the reference cannot be run here and ships no golden vectors.
 *
 * Everything here is written so that every floating-point operation is a single
 * IEEE-754 binary64 operation (compile with -ffp-contract=off): the CUDA engine
 * evaluates the same operation sequence with __dmul_rn/__dadd_rn/__ddiv_rn.
 */
#ifndef ORACLE_MATH_H
#define ORACLE_MATH_H

#include <stdint.h>
#include <string.h>

#define OR_GOLDEN 0x9E3779B97F4A7C15ULL
#define OR_K1     0xBF58476D1CE4E5B9ULL
#define OR_K2     0x94D049BB133111EBULL
#define OR_KCTR   0xD1B54A32D192ED03ULL

static inline uint64_t or_bits(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double or_from_bits(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

/* 64-bit finaliser (the splitmix64 output function). */
static inline uint64_t or_fin(uint64_t z)
{
    z ^= z >> 30; z *= OR_K1;
    z ^= z >> 27; z *= OR_K2;
    z ^= z >> 31;
    return z;
}

/* Per-LP stream key and the c-th draw of LP i (stateless, counter based). */
static inline uint64_t or_lpkey(uint64_t seed, uint32_t lp)
{
    return or_fin(seed ^ (((uint64_t)lp + 1ULL) * OR_GOLDEN));
}
static inline uint64_t or_draw(uint64_t lpkey, uint64_t ctr)
{
    return or_fin(lpkey + ctr * OR_KCTR);
}

/* Uniform on the open interval (0,1): 52 random bits plus one half, exact in binary64. */
static inline double or_u01(uint64_t d)
{
    return ((double)(d >> 12) + 0.5) * 0x1p-52;
}

/* floor(d * n / 2^64) */
static inline uint64_t or_mulhi(uint64_t d, uint64_t n)
{
    return (uint64_t)(((unsigned __int128)d * (unsigned __int128)n) >> 64);
}

/*
 * Natural logarithm of a positive, finite, normal binary64 number (0 gives -inf).
 * Classic argument reduction x = 2^k (1+f), sqrt(1/2) < 1+f <= sqrt(2),
 * s = f/(2+f), log(1+f) = f - f^2/2 + s (f^2/2 + R(s^2)) with a degree-14 minimax R
 * (the coefficients below are the well-known fdlibm set), evaluated in one fixed
 * operation order so that CPU and GPU agree to the last bit.
 */
static inline double or_log(double x)
{
    static const double ln2_hi = 6.93147180369123816490e-01;
    static const double ln2_lo = 1.90821492927058770002e-10;
    static const double L1 = 6.666666666666735130e-01;
    static const double L2 = 3.999999999940941908e-01;
    static const double L3 = 2.857142874366239149e-01;
    static const double L4 = 2.222219843214978396e-01;
    static const double L5 = 1.818357216161805012e-01;
    static const double L6 = 1.531383769920937332e-01;
    static const double L7 = 1.479819860511658591e-01;
    uint64_t b = or_bits(x);
    if (b == 0) return -__builtin_inf();
    int64_t k = (int64_t)(b >> 52) - 1023;
    uint64_t m = b & 0x000FFFFFFFFFFFFFULL;
    /* mantissa above sqrt(2): use x/2 and bump the exponent */
    if (m > 0x6A09E667F3BCCULL) { k += 1; b = m | 0x3FE0000000000000ULL; }
    else                        {         b = m | 0x3FF0000000000000ULL; }
    double f  = or_from_bits(b) - 1.0;
    double dk = (double)k;
    double s  = f / (2.0 + f);
    double z  = s * s;
    double w  = z * z;
    double t1 = w * (L2 + w * (L4 + w * L6));
    double t2 = z * (L1 + w * (L3 + w * (L5 + w * L7)));
    double R  = t2 + t1;
    double hfsq = (0.5 * f) * f;
    return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
}

/* Exponential variate with mean mu from a uniform x: Haskell's `- log x * mu`. */
static inline double or_expo(double x, double mu)
{
    return -(or_log(x) * mu);
}

/* Order-dependent committed-trace hash: fold one executed event into h. */
static inline uint64_t or_trace(uint64_t h, double recv_ts, uint32_t src, uint64_t ctr)
{
    uint64_t x = or_bits(recv_ts) ^ (ctr * OR_K2) ^ ((uint64_t)src * OR_K1);
    return or_fin((h ^ x) + OR_GOLDEN);
}

/* ---- StdGen, `random` 1.1 regime (L'Ecuyer 1988 combined MLCG) ---- */
typedef struct { int32_t s1, s2; } or_stdgen;

/* mkStdGen :: Int -> StdGen  (mkStdGen32 of the low 32 bits, sign cleared) */
static inline or_stdgen or_mkstdgen(int64_t seed)
{
    int32_t s = (int32_t)((uint32_t)seed & 0x7FFFFFFFu);
    int32_t q = s / 2147483562, s1 = s % 2147483562;     /* s >= 0: divMod == quotRem */
    int32_t s2 = q % 2147483398;
    or_stdgen g = { s1 + 1, s2 + 1 };
    return g;
}

/* next :: StdGen -> (Int, StdGen), values in [1, 2147483562] */
static inline int32_t or_stdnext(or_stdgen *g)
{
    int32_t k  = g->s1 / 53668;
    int32_t a  = 40014 * (g->s1 - k * 53668) - k * 12211;
    if (a < 0) a += 2147483563;
    int32_t k2 = g->s2 / 52774;
    int32_t b  = 40692 * (g->s2 - k2 * 52774) - k2 * 3791;
    if (b < 0) b += 2147483399;
    int32_t z = a - b;
    if (z < 1) z += 2147483562;
    g->s1 = a; g->s2 = b;
    return z;
}

/*
 * random :: StdGen -> (Double, StdGen): an Int64 drawn by randomIvalInteger over the
 * full Int64 range (base b = 2147483562, three `next` calls until the magnitude
 * exceeds 2^64 * 1000), masked to 53 bits and divided by 2^53.  Only the value
 * modulo 2^53 matters, so wrapping 64-bit arithmetic is exact.
 */
static inline double or_stdrandom_double(or_stdgen *g)
{
    const uint64_t b = 2147483562ULL;
    uint64_t v = 0;
    for (int i = 0; i < 3; ++i) {
        uint64_t x = (uint64_t)or_stdnext(g) - 1ULL;
        v = v * b + x;
    }
    return (double)(v & ((1ULL << 53) - 1ULL)) * 0x1p-53;
}

#endif
