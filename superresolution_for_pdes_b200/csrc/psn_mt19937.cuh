// psn_mt19937.cuh -- bit-exact replay of numpy's legacy global RNG (MT19937) on the device.
//
// The reference draws every variable-coefficient field on the host from the global legacy stream:
//     theta = np.random.uniform(0.5, 2.0, (n, n))      src/resolution_comparison.py:42, src/subdomain_upscaling.py:31,
//                                                      src/compare_test_cases.py:39
// preceded per example by the two wave-number draws np.random.uniform(*k_range) (src/resolution_comparison.py:27-28).
// At 640 x 640 that is 2 ms of host time per example -- two orders of magnitude more than the solve it feeds -- and
// 3.3 MB that would have to cross PCIe.  Here the SAME numbers are produced on the device:
//   * numpy's uniform is  low + (high - low) * d,  d = ((a >> 5) * 2^26 + (b >> 6)) / 2^53  with a, b two successive
//     tempered MT19937 words (numpy/random/src/mt19937/mt19937.h: mt19937_next_double; _legacy/distributions:
//     random_uniform) -- reproduced with explicit __dmul_rn / __dadd_rn (no FMA contraction, as in numpy's x86-64 build);
//   * the stream is split into independent blocks (one per example) by JUMP-AHEAD: the state after J more words is
//     g(F) s with g(t) = t^J mod chi(t), F the one-word transition and chi its characteristic polynomial (degree
//     19937).  chi is obtained once per process with Berlekamp-Massey from the generator's own output, g by
//     square-and-multiply on the host (cached per J); g(F) s is evaluated on the device as a binary convolution:
//     (g(F) s)[j] = XOR of x[i + j] over the set bits i of g, x the source state's own output sequence extended by
//     ~20,000 words in shared memory -- no dependent steps, so 16 CTAs share one jump;
//   * block b's state is reached in ceil(log2 B) rounds of doubling (round k: states 2^k .. 2^(k+1)-1 from states
//     0 .. 2^k-1 with the polynomial of J * 2^k), every round one launch with one CTA per state;
//   * one CTA per block then regenerates ("twists") its state in shared memory -- three barrier-separated phases of
//     227 / 227 / 170 independent words -- tempers and writes 312 doubles per twist.
// Host logic (polynomials, and psn_mt19937_jump_host for the tests) is plain C++; the stream itself is only ever
// produced by the kernels below.
#pragma once
#include <stdint.h>
#include <string.h>
#include <map>
#include <vector>

namespace psn_mt {

constexpr int N = 624, M = 397;
constexpr uint32_t MATRIX_A = 0x9908b0dfu, UPPER = 0x80000000u, LOWER = 0x7fffffffu;
constexpr int DEG = 19937;
constexpr int PW = 313;          // 64-bit words of a polynomial of degree <= 19937 (+ slack)
constexpr int CHUNKS = 624;      // 32-bit chunks of a jump polynomial (degree < 19968)
constexpr int QB = 32;           // Horner chunk width (bits)

#if defined(__CUDACC__)
#define PSN_MT_HD __host__ __device__ __forceinline__
#else
#define PSN_MT_HD inline
#endif

PSN_MT_HD uint32_t twist(uint32_t hi, uint32_t lo, uint32_t far) {
    const uint32_t y = (hi & UPPER) | (lo & LOWER);
    return far ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
}
PSN_MT_HD uint32_t temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

// ----------------------------------------------------------------------------- host: polynomials over GF(2)
struct Poly {
    uint64_t w[PW];
    Poly() { memset(w, 0, sizeof(w)); }
    bool bit(int i) const { return (w[i >> 6] >> (i & 63)) & 1u; }
    void flip(int i) { w[i >> 6] ^= (uint64_t)1 << (i & 63); }
};

// one-word step of the generator on a plain (non-circular) window, host side: returns x[i + 624]
inline void host_step(std::vector<uint32_t>& win, size_t& head) {
    const size_t n = N;
    const uint32_t nw = twist(win[head], win[(head + 1) % n], win[(head + M) % n]);
    win[head] = nw;
    head = (head + 1) % n;
}

// characteristic polynomial of the one-word transition: Berlekamp-Massey on 2 * 19937 + 64 output bits
inline const Poly& chi() {
    static Poly c;
    static bool done = false;
    if (done) return c;
    const int L2 = 2 * DEG + 64;
    std::vector<uint32_t> win(N);
    uint32_t s = 19650218u;   // any non-zero state
    for (int i = 0; i < N; ++i) {
        win[i] = s;
        s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)(i + 1);
    }
    std::vector<uint8_t> seq(L2);
    size_t head = 0;
    for (int i = 0; i < L2; ++i) {
        seq[i] = (uint8_t)(win[(head + 1) % N] & 1u);   // a bit of the live state (word 0 only contributes its top bit)
        host_step(win, head);
    }
    // C(x): connection polynomial, s_n = sum_{i=1..L} C_i s_{n-i}
    const int W = (DEG + 64 + 63) / 64 + 1;
    std::vector<uint64_t> C(W, 0), B(W, 0), T(W, 0), R(W, 0);
    C[0] = B[0] = 1;
    int L = 0, m = 1;
    for (int n = 0; n < L2; ++n) {
        // R bit i = s[n - i]
        for (int k = W - 1; k > 0; --k) R[k] = (R[k] << 1) | (R[k - 1] >> 63);
        R[0] = (R[0] << 1) | seq[n];
        uint64_t acc = 0;
        const int lw = L / 64 + 1;
        for (int k = 0; k < lw && k < W; ++k) acc ^= C[k] & R[k];
        const int d = __builtin_parityll(acc);
        if (!d) {
            ++m;
            continue;
        }
        const bool grow = 2 * L <= n;
        if (grow) T = C;
        // C ^= B << m
        const int ws = m / 64, bs = m % 64;
        for (int k = W - 1; k >= ws; --k) {
            uint64_t v = B[k - ws] << bs;
            if (bs && k - ws - 1 >= 0) v |= B[k - ws - 1] >> (64 - bs);
            C[k] ^= v;
        }
        if (grow) {
            L = n + 1 - L;
            B = T;
            m = 1;
        } else {
            ++m;
        }
    }
    // chi(t) = t^L C(1/t): coefficient i of chi = C_{L-i}
    if (L == DEG) {
        for (int i = 0; i <= DEG; ++i)
            if ((C[(L - i) >> 6] >> ((L - i) & 63)) & 1u) c.flip(i);
    }
    done = true;
    return c;
}
inline bool chi_ok() { return chi().bit(DEG); }

// r = a * a mod chi
inline void sqr_mod(const Poly& a, Poly& r) {
    static const Poly& X = chi();
    const int W2 = 2 * PW;
    std::vector<uint64_t> t(W2 + 1, 0);
    for (int i = 0; i < DEG; ++i)
        if (a.bit(i)) t[(2 * i) >> 6] |= (uint64_t)1 << ((2 * i) & 63);
    for (int i = 2 * DEG - 2; i >= DEG; --i) {
        if (!((t[i >> 6] >> (i & 63)) & 1u)) continue;
        const int sh = i - DEG, ws = sh >> 6, bs = sh & 63;
        for (int k = 0; k < PW; ++k) {
            t[k + ws] ^= X.w[k] << bs;
            if (bs) t[k + ws + 1] ^= X.w[k] >> (64 - bs);
        }
    }
    for (int k = 0; k < PW; ++k) r.w[k] = t[k];
}
// r = r * t mod chi
inline void mulx_mod(Poly& r) {
    static const Poly& X = chi();
    for (int k = PW - 1; k > 0; --k) r.w[k] = (r.w[k] << 1) | (r.w[k - 1] >> 63);
    r.w[0] <<= 1;
    if (r.bit(DEG))
        for (int k = 0; k < PW; ++k) r.w[k] ^= X.w[k];
}
// t^e mod chi
inline Poly pow_t(uint64_t e) {
    Poly r;
    r.w[0] = 1;
    for (int b = 63; b >= 0; --b) {
        Poly q;
        sqr_mod(r, q);
        r = q;
        if ((e >> b) & 1u) mulx_mod(r);
    }
    return r;
}
// 32-bit chunks (low first) of a polynomial of degree < 19968
inline void to_chunks(const Poly& p, uint32_t* out) {
    for (int c = 0; c < CHUNKS; ++c) out[c] = (uint32_t)(p.w[c >> 1] >> ((c & 1) * 32));
}

// cache: jump polynomials t^(J 2^k) mod chi, k = 0 .. (grown on demand by squaring)
struct JumpFamily {
    std::vector<Poly> g;
};
inline const Poly& jump_poly(uint64_t J, int k) {
    static std::map<uint64_t, JumpFamily> cache;
    JumpFamily& f = cache[J];
    if (f.g.empty()) f.g.push_back(pow_t(J));
    while ((int)f.g.size() <= k) {
        Poly q;
        sqr_mod(f.g.back(), q);
        f.g.push_back(q);
    }
    return f.g[k];
}

// g(F) s on the host (Horner, one bit at a time): tests and small cases only
inline void host_apply(const Poly& g, const uint32_t* s, uint32_t* out) {
    std::vector<uint32_t> acc(N, 0);
    size_t head = 0;
    int top = DEG - 1;
    while (top > 0 && !g.bit(top)) --top;
    for (int i = top; i >= 0; --i) {
        host_step(acc, head);
        if (g.bit(i))
            for (int j = 0; j < N; ++j) acc[(head + j) % N] ^= s[j];
    }
    for (int j = 0; j < N; ++j) out[j] = acc[(head + j) % N];
}

}  // namespace psn_mt

#if defined(__CUDACC__)
// ----------------------------------------------------------------------------- device kernels
#define PSN_MT_THREADS 640   // 20 warps: one thread per state word (624 active)
#define PSN_MT_SPLIT 16      // CTAs that share one jump (each takes 39 of the 624 polynomial chunks)
#define PSN_MT_SEQ (psn_mt::N + 88 * 227)   // extended source sequence x[0 .. 20599] (x[19936 + 623] is the last one used)

// states[dst0 + c] ^= (part of g)(F) states[src0 + c] for c = blockIdx.x (part = blockIdx.y); the destination states must be zero before
// the launch.  g(F) s = sum_i g_i F^i s and (F^i s)[j] = x[i + j]: with the source sequence extended to
// x[0 .. 19936 + 623] in shared memory (88 barrier-separated rounds of 227 independent words) output word j is the XOR of
// x[i + j] over the set bits i of g -- no dependent steps, so PSN_MT_SPLIT CTAs take disjoint bit ranges of g and
// combine with atomicXor.  (A Horner formulation with one CTA per jump took 0.6 ms per doubling round, 6 ms for 512
// blocks; this one is bound by the ~10^4 shared-memory loads per thread of a single round.)
__global__ void __launch_bounds__(PSN_MT_THREADS) psn_mt_jump_kernel(uint32_t* states, long long src0, long long dst0,
                                                                     const uint32_t* __restrict__ gchunks) {
    using namespace psn_mt;
    extern __shared__ uint32_t mt_smem[];
    uint32_t* X = mt_smem;               // [PSN_MT_SEQ]
    uint32_t* G = mt_smem + PSN_MT_SEQ;  // [CHUNKS / PSN_MT_SPLIT]
    const int t = threadIdx.x;
    constexpr int CPB = CHUNKS / PSN_MT_SPLIT;
    const int c0 = blockIdx.y * CPB;
    const uint32_t* src = states + (src0 + blockIdx.x) * (long long)N;
    if (t < N) X[t] = src[t];
    if (t < CPB) G[t] = gchunks[c0 + t];
    __syncthreads();
    // only the part of the sequence this CTA reads: x[32 c0 .. 32 (c0 + CPB) + 623]
    const int need = 32 * (c0 + CPB) + N;
    for (int base = N; base < need; base += 227) {
        if (t < 227) X[base + t] = twist(X[base + t - N], X[base + t - N + 1], X[base + t - 227]);
        __syncthreads();
    }
    if (t < N) {
        uint32_t a = 0u;
        for (int c = 0; c < CPB; ++c) {
            uint32_t v = G[c];          // uniform
            const uint32_t* xp = X + 32 * (c0 + c) + t;
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if ((v >> i) & 1u) a ^= xp[i];
        }
        if (a) atomicXor(states + (dst0 + blockIdx.x) * (long long)N + t, a);
    }
}
inline size_t psn_mt_jump_smem() { return (size_t)(PSN_MT_SEQ + psn_mt::CHUNKS / PSN_MT_SPLIT) * sizeof(uint32_t); }

// One CTA per block of the stream: emits `skip` raw doubles in [0, 1) followed by `count` doubles low + range * d,
// starting at word `pos` of its state.
#define PSN_MT_GEN_THREADS 320
__global__ void __launch_bounds__(PSN_MT_GEN_THREADS) psn_mt_uniform_kernel(const uint32_t* __restrict__ states, int pos,
                                                                            long long skip, long long count, double low, double range,
                                                                            double* __restrict__ out, double* __restrict__ skipped) {
    using namespace psn_mt;
    __shared__ uint32_t K[2][N];
    __shared__ uint32_t carry;
    const int t = threadIdx.x;
    const long long blk = blockIdx.x;
    const uint32_t* src = states + blk * (long long)N;
    for (int i = t; i < N; i += PSN_MT_GEN_THREADS) K[0][i] = src[i];
    __syncthreads();
    int cur = 0;
    const long long total_words = 2 * (skip + count);
    long long q0 = 0;          // stream index (in words, relative to the block's first word) of K[cur][first]
    int first = pos;           // first unused word of the current key (624: none)
    double* outb = out + blk * count;
    double* skb = skipped ? skipped + blk * skip : nullptr;
    while (q0 < total_words) {
        if (first >= N) {
            // regenerate: K[nxt] from K[cur]
            const int nxt = cur ^ 1;
            const uint32_t* a = K[cur];
            uint32_t* b = K[nxt];
            if (t < 227) b[t] = twist(a[t], a[t + 1], a[t + M]);
            __syncthreads();
            if (t < 227) b[227 + t] = twist(a[227 + t], a[228 + t], b[t]);
            __syncthreads();
            if (t < 170) {
                const int i = 454 + t;
                const uint32_t lo = (i == N - 1) ? b[0] : a[i + 1];
                b[i] = twist(a[i], lo, b[i - 227]);
            }
            __syncthreads();
            cur = nxt;
            first = 0;
        }
        const uint32_t* k = K[cur];
        long long avail = N - first;
        if (avail > total_words - q0) avail = total_words - q0;
        // doubles whose SECOND word lies in [q0, q0 + avail): d = ceil((q0 - 1) / 2) ... ; first word may be the carry
        const long long d_lo = q0 >> 1;                       // first double with word 2d+1 >= q0 (2d+1 >= q0 <=> d >= (q0-1)/2 rounded up = q0 >> 1)
        const long long d_hi = (q0 + avail) >> 1;             // doubles d < d_hi have word 2d+1 < q0 + avail
        for (long long d = d_lo + t; d < d_hi; d += PSN_MT_GEN_THREADS) {
            const long long wa = 2 * d - q0;                  // index of the first word relative to `first` (-1: carry)
            const uint32_t a32 = (wa < 0) ? carry : temper(k[first + wa]);
            const uint32_t b32 = temper(k[first + wa + 1]);
            const double dv = ((double)(a32 >> 5) * 67108864.0 + (double)(b32 >> 6)) * (1.0 / 9007199254740992.0);
            if (d < skip) {
                if (skb) skb[d] = dv;
            } else {
                outb[d - skip] = __dadd_rn(low, __dmul_rn(range, dv));
            }
        }
        __syncthreads();    // everyone has read the carry of the previous round
        if (((q0 + avail) & 1) && t == 0) carry = temper(k[first + avail - 1]);
        q0 += avail;
        first += (int)avail;
        __syncthreads();
    }
}
#endif  // __CUDACC__
