/*
 * prove_division.c -- computer-assisted proof that the one-step shared-reciprocal quotient used by
 * the extend kernel (uvrt_kernels.cuh, DIV_MARKSTEIN1),
 *
 *     r   = RN(1/d)               (once per ray and axis, __frcp_rn)
 *     q0  = RN(n * r)
 *     rem = n - d*q0              (one FMA, exact)
 *     q1  = RN(q0 + rem * r)      (one FMA)
 *
 * equals the IEEE-754 quotient RN(n/d) for ALL binary32 n, d whose intermediates stay in the
 * normal range (the kernel's ray_is_tame() / scene check guarantee that).
 *
 * Argument ("Division"): write n = a*2^x, d = b*2^y with integer significands
 * a, b in [2^23, 2^24).  The value before the last rounding is n/d + rem*delta with
 * delta = r - 1/d, and |rem*delta| is below the distance from n/d to the nearest rounding
 * boundary (a midpoint of two adjacent floats) unless that distance is unusually small:
 *     a >= b : midpoints are (2k+1)*2^-24;  n/d - m = (a*2^24 - (2k+1)*b) / (b*2^24)
 *     a <  b : midpoints are (2k+1)*2^-25;  n/d - m = (a*2^25 - (2k+1)*b) / (b*2^25)
 * The error term can only reach the boundary when the integer numerator |num| is at most 1
 * (a >= b) or 2 (a < b).  Those hard cases are finite: for every b and every small num the
 * congruence a * 2^K = num (mod b) pins a down.  This program enumerates them -- generously, all
 * |num| <= 8, every b, both significand orderings, both signs, several exponents -- and checks
 * the sequence against the hardware's correctly rounded division.  It also checks the two-step
 * form and 2^32 random operand pairs.
 *
 * Build: gcc -O2 -mfma -fopenmp -ffp-contract=off -o prove_division prove_division.c -lm
 * Exit status 0 and "FAILURES 0" mean the claim holds.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline float one_step(float n, float d, float r)
{
    float q = n * r;
    float rem = fmaf(-d, q, n);
    return fmaf(rem, r, q);
}
static inline float two_step(float n, float d, float r)
{
    float q = one_step(n, d, r);
    float rem = fmaf(-d, q, n);
    return fmaf(rem, r, q);
}

/* modular inverse of x modulo m (m odd or gcd(x,m)=1), 64-bit */
static int64_t inv_mod(int64_t x, int64_t m)
{
    int64_t a = x % m, b = m, u = 1, v = 0;
    while (b) {
        int64_t t = a / b;
        a -= t * b; { int64_t s = a; a = b; b = s; }
        u -= t * v; { int64_t s = u; u = v; v = s; }
    }
    if (a != 1) return -1;
    u %= m;
    if (u < 0) u += m;
    return u;
}

static int check(int64_t a, int64_t b, unsigned long long* tested)
{
    /* exponent pairs (n, d): includes the smallest numerators the kernel admits (denormal residuals) */
    static const int ex[6][2] = {{0, 0}, {-43, 0}, {5, -30}, {-20, -17}, {-100, 0}, {-100, -30}};
    int bad = 0;
    for (int e = 0; e < 6; e++)
        for (int s = 0; s < 4; s++) {
            float n = ldexpf((float)a, ex[e][0] - 23), d = ldexpf((float)b, ex[e][1] - 23);
            if (s & 1) n = -n;
            if (s & 2) d = -d;
            float r = 1.0f / d, q = n / d;
            float q1 = one_step(n, d, r), q2 = two_step(n, d, r);
            if (q1 != q || q2 != q) {
                bad++;
                if (bad < 4) printf("MISMATCH n=%a d=%a exact=%a one=%a two=%a\n", n, d, q, q1, q2);
            }
            (*tested)++;
        }
    return bad;
}

int main(int argc, char** argv)
{
    const int randomLog2 = argc > 1 ? atoi(argv[1]) : 32;   /* 2^randomLog2 random pairs */
    const int64_t LO = 1 << 23, HI = 1 << 24;
    unsigned long long failures = 0, tested = 0, hard = 0;
#pragma omp parallel for schedule(dynamic, 4096) reduction(+ : failures, tested, hard)
    for (int64_t b = LO; b < HI; b++) {
        for (int K = 24; K <= 25; K++) {               /* K = 24: a >= b ;  K = 25: a < b */
            const int64_t P2 = (int64_t)1 << K;
            int v = 0;
            while (v < K && !((b >> v) & 1)) v++;
            const int64_t g = (int64_t)1 << v;         /* gcd(2^K, b) */
            const int64_t bm = b / g;
            const int64_t iv = bm == 1 ? 0 : inv_mod((P2 / g) % bm, bm);
            for (int64_t num = -8; num <= 8; num++) {
                if (num == 0 || (num % g) != 0) continue;
                /* a * (2^K/g) = num/g  (mod b/g) */
                int64_t t = (num / g) % bm;
                if (t < 0) t += bm;
                int64_t a0 = bm == 1 ? 0 : (int64_t)(((__int128)t * iv) % bm);
                int64_t from = K == 24 ? b : LO, to = K == 24 ? HI : b;   /* range of a */
                int64_t first = a0 + ((from - a0 + bm - 1) / bm) * bm;
                for (int64_t a = first; a < to; a += bm) {
                    /* the quotient (a*2^K - num)/b must be an odd integer: a midpoint */
                    __int128 top = (__int128)a * P2 - num;
                    if (top % b != 0) continue;
                    int64_t M = (int64_t)(top / b);
                    if (!(M & 1)) continue;
                    hard++;
                    failures += check(a, b, &tested);
                }
            }
        }
    }
    printf("hard cases (|num| <= 8) %llu, operand pairs tested %llu\n", hard, tested);

    /* random and adversarial pairs over the exponent ranges the kernel admits */
    unsigned long long rnd = 0;
#pragma omp parallel reduction(+ : failures, rnd)
    {
        uint32_t s = 0x9e3779b9u;
#ifdef _OPENMP
        extern int omp_get_thread_num(void);
        s *= (uint32_t)(omp_get_thread_num() * 2 + 1);
#endif
#pragma omp for
        for (long long i = 0; i < (1ll << randomLog2); i++) {
            s ^= s << 13; s ^= s >> 17; s ^= s << 5; uint32_t md = s & 0x7fffff;
            s ^= s << 13; s ^= s >> 17; s ^= s << 5; uint32_t mn = s & 0x7fffff;
            s ^= s << 13; s ^= s >> 17; s ^= s << 5; uint32_t e = s;
            int ed = -(int)(e % 31), en = -100 + (int)((e >> 8) % 121);
            if ((e >> 20) & 1) md = ((e >> 21) & 1) ? 0x7fffff - (md & 0xff) : (md & 0xff);
            if ((e >> 22) & 1) mn = ((e >> 23) & 1) ? 0x7fffff - (mn & 0xff) : (mn & 0xff);
            uint32_t du = ((uint32_t)(ed + 127) << 23) | md | (((e >> 30) & 1) << 31);
            uint32_t nu = ((uint32_t)(en + 127) << 23) | mn | ((e >> 31) << 31);
            float d, n;
            memcpy(&d, &du, 4);
            memcpy(&n, &nu, 4);
            float r = 1.0f / d, q = n / d;
            failures += (one_step(n, d, r) != q) + (two_step(n, d, r) != q);
            rnd++;
        }
    }
    printf("random pairs tested %llu\n", rnd);
    printf("FAILURES %llu\n", failures);
    return failures ? 1 : 0;
}
