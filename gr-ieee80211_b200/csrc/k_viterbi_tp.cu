// k_viterbi_tp.cu -- decode block, throughput variant: ONE THREAD PER FRAME.
//
// Same function and same results as k_viterbi.cu (lib/decode_impl.cc:164-520: depuncture, 64-state soft
// Viterbi with float32 metrics in the reference's operation order, full traceback from state 0, descramble,
// A-MPDU walk, CRC-32), organised for large batches: a thread keeps all 64 path metrics of its frame in
// registers, so a trellis step is 32 butterflies of straight-line code with NO shuffles and NO shared-memory
// traffic in the dependent chain -- the branch-metric class of every butterfly is a compile-time constant
// (9 instructions per butterfly: <=4 FADD, 2 FADD for the decision signs, 2 FMNMX, 2 SHF), 32 independent
// chains per thread.  k_viterbi.cu (one warp per frame pair, 15 instr per lane and step behind a ~100-cycle
// shuffle chain) remains the low-latency path for small batches; ctx.cu picks by batch size.
//
// Data movement.  A warp owns 32 frames.  Soft bits are staged per 30-step chunk: for each of its frames the
// warp reads the chunk's soft bits coalesced, depunctures them (lib/cloud80211phy.cc:1857-1860 patterns in
// closed form) and writes (t0,t1) pairs into a shared-memory row per frame; lane j then walks row j.
// Survivors: one 64-bit decision word per thread and step, stored [step][thread] so a warp writes 256 B
// contiguous; the traceback reads them back the same way.  Decoded bits go to a [word][thread] scratch,
// are descrambled there and walked by the same thread for the A-MPDU / CRC-32 / PDU output.
#include "common.cuh"

namespace {

constexpr int TPB = 128;                       // threads (= frames) per CTA (224 x 2 CTAs = 14 warps per SM at 144 registers spills: slower)
constexpr int CS = 30;                         // trellis steps per staged chunk (multiple of every puncture period and of 2)
constexpr int ROWF2 = CS + 1;                  // float2 per shared-memory row (odd -> conflict-free 8-byte column walks)

// encoder output class of the transition 2k --0--> k (same definition as lut.cc bmClass), compile time
__host__ __device__ constexpr int parity7(int v) { return (v ^ (v >> 1) ^ (v >> 2) ^ (v >> 3) ^ (v >> 4) ^ (v >> 5) ^ (v >> 6)) & 1; }
__host__ __device__ constexpr int bm_class(int k)
{
    int s6 = 2 * k, reg = 0;
    for (int q = 0; q < 6; q++) reg |= ((s6 >> (5 - q)) & 1) << (q + 1);
    return parity7(reg & 0155) * 2 + parity7(reg & 0117);
}

// one trellis step: m (metrics of the 64 old states) -> n.  wlo/whi: decision bits of new states 0..31 / 32..63.
// tab = {0, t1, t0, t1+t0}; pre + tab[0] is pre itself (adding +0.0f is exact), so those adds are skipped.
template <int K>
__device__ __forceinline__ void butterfly(const float (&m)[64], float (&n)[64], const float t0, const float t1, const float t3,
                                          uint32_t& wlo, uint32_t& whi)
{
    constexpr int c = bm_class(K);
    const float e = m[2 * K], o = m[2 * K + 1];
    float eLo, oLo, eHi, oHi;                                       // A = tab[c] on (e->lo, o->hi), B = tab[3-c] on (o->lo, e->hi)
    if (c == 0) { eLo = e; oLo = __fadd_rn(o, t3); eHi = __fadd_rn(e, t3); oHi = o; }
    else if (c == 1) { eLo = __fadd_rn(e, t1); oLo = __fadd_rn(o, t0); eHi = __fadd_rn(e, t0); oHi = __fadd_rn(o, t1); }
    else if (c == 2) { eLo = __fadd_rn(e, t0); oLo = __fadd_rn(o, t1); eHi = __fadd_rn(e, t1); oHi = __fadd_rn(o, t0); }
    else { eLo = __fadd_rn(e, t3); oLo = o; eHi = e; oHi = __fadd_rn(o, t3); }
    // odd predecessor wins only if strictly larger  <=>  (even - odd) is negative (ties keep the even one)
    wlo = __funnelshift_l(__float_as_uint(__fsub_rn(eLo, oLo)), wlo, 1);
    whi = __funnelshift_l(__float_as_uint(__fsub_rn(eHi, oHi)), whi, 1);
    n[K] = fmaxf(eLo, oLo);
    n[K + 32] = fmaxf(eHi, oHi);
}

template <int K>
struct Bf {
    static __device__ __forceinline__ void run(const float (&m)[64], float (&n)[64], float t0, float t1, float t3, uint32_t& wlo, uint32_t& whi)
    {
        butterfly<K>(m, n, t0, t1, t3, wlo, whi);                   // K = 31 first: its bit ends up at position 31
        Bf<K - 1>::run(m, n, t0, t1, t3, wlo, whi);
    }
};
template <>
struct Bf<-1> {
    static __device__ __forceinline__ void run(const float (&)[64], float (&)[64], float, float, float, uint32_t&, uint32_t&) {}
};

__device__ __forceinline__ uint2 acs(const float (&m)[64], float (&n)[64], const float2 tt)
{
    uint32_t wlo = 0, whi = 0;
    Bf<31>::run(m, n, tt.x, tt.y, __fadd_rn(tt.y, tt.x), wlo, whi);
    return make_uint2(wlo, whi);
}

// ---------------------------------------------------------------------------------------------------------------
// Packed forward pass (C8B_TP_PACKED = 1, NOT the default): the same add-compare-select in the same float32 operation order, with the
// adds and the decision differences issued as sm_100 two-wide FADD2 (add.rn.f32x2: two independent IEEE adds, so every
// sum is the bit the scalar form produces).  The kernel is bound by instruction issue, not by the FP32 pipe, so halving the
// FADD count is what pays: 14 instructions per butterfly PAIR (4 FADD2 adds, 2 FADD2 differences, 4 FMNMX, 4 SHF) instead
// of 18.
//
// The 64 metrics live in 32 aligned register pairs.  Layout L_q pairs the two states that differ in bit q.  With the input
// in L_q (q >= 1) the butterflies k and k' = k | 1 << (q-1) read their even predecessors from ONE pair (2k, 2k') and their
// odd predecessors from another (2k+1, 2k'+1), and their results (k, k') and (k+32, k'+32) are pairs of L_(q-1): the
// layout walks 5 -> 4 -> 3 -> 2 -> 1 -> 0 for free.  In L_0 a pair holds both predecessors of one butterfly; that step
// adds (m_even, m_odd) + (A, B) and + (B, A), compares inside the pairs and writes (k, k+32) = a pair of L_5 again.
// Period 6 divides the 30-step staging chunk.  Branch-metric pairs: the code is linear, so class(k') = class(k) ^ delta_q
// and a step needs just the four pairs (tab[x], tab[x ^ delta_q]).
// MEASURED (one B200, 56832-frame wave, profiles/ncu_vtp_packed_r01.md): bit-exact (all decode / chain tests, 1,048,486 MPDUs of
// the bench identical) and 242 instead of 290 instructions per trellis step, but SLOWER: 9.3 ms against 8.5 ms for the scalar
// butterflies.  FADD2 issues only to the FMA-heavy sub-pipe and reads / writes 64-bit operands: issue-slot utilisation drops from
// 75 % to 58 % behind dispatch stalls (0.57 per issue) and math-pipe throttle (0.78), and the 23 KB loop body adds
// instruction-fetch stalls (0.32).  Kept behind the knob as the record of that experiment.
// Decision words keep their natural bit order (state n -> bit n & 31 of word n >> 5), so survivors and traceback are
// unchanged: q <= 3 shifts the sign bits in per group of 2^q butterflies, q = 4 / 5 fill two half-words and merge them with
// one PRMT / one shift-or per word.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 swap2(u64 v) { float lo, hi; upk2(v, lo, hi); return pk2(hi, lo); }   // free: FADD2 operand swizzle .LO_HI
__device__ __forceinline__ u64 max2(u64 a, u64 b)
{
    float a0, a1, b0, b1;
    upk2(a, a0, a1); upk2(b, b0, b1);
    return pk2(fmaxf(a0, b0), fmaxf(a1, b1));
}
__device__ __forceinline__ uint32_t push_sign(float d, uint32_t w) { return __funnelshift_l(__float_as_uint(d), w, 1); }

__host__ __device__ constexpr int rm_bit(int s, int q) { return ((s >> (q + 1)) << q) | (s & ((1 << q) - 1)); }
__host__ __device__ constexpr int ins_bit(int p, int q, int b) { return ((p >> q) << (q + 1)) | (b << q) | (p & ((1 << q) - 1)); }

// butterflies k = ins_bit(J, Q-1, 0) and k | 1 << (Q-1) of a step whose input is in layout L_Q; TP[x] = (tab[x], tab[x ^ delta_Q])
template <int Q, int J>
__device__ __forceinline__ void bf_pair(const u64 (&P)[32], u64 (&N)[32], const u64 (&TP)[4], u64& dLo, u64& dHi)
{
    constexpr int k = ins_bit(J, Q - 1, 0);
    constexpr int c = bm_class(k);
    constexpr bool zeroA = c == 0 && bm_class(1 << (Q - 1)) == 0;       // (0, 0): adding it is the identity
    constexpr bool zeroB = (c ^ 3) == 0 && bm_class(1 << (Q - 1)) == 0;
    const u64 E = P[rm_bit(2 * k, Q)], O = P[rm_bit(2 * k + 1, Q)];
    const u64 aLo = zeroA ? E : add2(E, TP[c]), bLo = zeroB ? O : add2(O, TP[c ^ 3]);
    const u64 aHi = zeroB ? E : add2(E, TP[c ^ 3]), bHi = zeroA ? O : add2(O, TP[c]);
    dLo = sub2(aLo, bLo);                                               // sign set <=> the odd predecessor is strictly larger
    dHi = sub2(aHi, bHi);
    N[rm_bit(k, Q - 1)] = max2(aLo, bLo);
    N[rm_bit(k + 32, Q - 1)] = max2(aHi, bHi);
}

// Q <= 3: groups of 2^Q butterflies, sign bits shifted in from the highest state down (natural order)
template <int Q, int G, int I>
struct GroupPairs {
    static __device__ __forceinline__ void run(const u64 (&P)[32], u64 (&N)[32], const u64 (&TP)[4], float (&dl)[1 << Q], float (&dh)[1 << Q])
    {
        constexpr int H = 1 << (Q - 1);
        u64 a, b;
        bf_pair<Q, G * H + I>(P, N, TP, a, b);
        upk2(a, dl[I], dl[I + H]);
        upk2(b, dh[I], dh[I + H]);
        GroupPairs<Q, G, I - 1>::run(P, N, TP, dl, dh);
    }
};
template <int Q, int G>
struct GroupPairs<Q, G, -1> {
    static __device__ __forceinline__ void run(const u64 (&)[32], u64 (&)[32], const u64 (&)[4], float (&)[1 << Q], float (&)[1 << Q]) {}
};
template <int Q, int G>
struct Groups {
    static __device__ __forceinline__ void run(const u64 (&P)[32], u64 (&N)[32], const u64 (&TP)[4], uint32_t& wlo, uint32_t& whi)
    {
        float dl[1 << Q], dh[1 << Q];
        GroupPairs<Q, G, (1 << (Q - 1)) - 1>::run(P, N, TP, dl, dh);
#pragma unroll
        for (int o = (1 << Q) - 1; o >= 0; o--) { wlo = push_sign(dl[o], wlo); whi = push_sign(dh[o], whi); }
        Groups<Q, G - 1>::run(P, N, TP, wlo, whi);
    }
};
template <int Q>
struct Groups<Q, -1> {
    static __device__ __forceinline__ void run(const u64 (&)[32], u64 (&)[32], const u64 (&)[4], uint32_t&, uint32_t&) {}
};

// Q = 4, 5: A collects the states with bit Q-1 set, B the others, both from the highest pair index down
template <int Q, int J>
struct Halves {
    static __device__ __forceinline__ void run(const u64 (&P)[32], u64 (&N)[32], const u64 (&TP)[4], uint32_t& aLo, uint32_t& bLo, uint32_t& aHi,
                                               uint32_t& bHi)
    {
        u64 a, b;
        float d0, d1;
        bf_pair<Q, J>(P, N, TP, a, b);
        upk2(a, d0, d1);
        bLo = push_sign(d0, bLo); aLo = push_sign(d1, aLo);
        upk2(b, d0, d1);
        bHi = push_sign(d0, bHi); aHi = push_sign(d1, aHi);
        Halves<Q, J - 1>::run(P, N, TP, aLo, bLo, aHi, bHi);
    }
};
template <int Q>
struct Halves<Q, -1> {
    static __device__ __forceinline__ void run(const u64 (&)[32], u64 (&)[32], const u64 (&)[4], uint32_t&, uint32_t&, uint32_t&, uint32_t&) {}
};

// Q = 0: pair K holds (m[2K], m[2K+1]); TQ[c] = (tab[c], tab[c ^ 3]); result pair K of L_5 = states (K, K + 32)
template <int K>
struct Inner {
    static __device__ __forceinline__ void run(const u64 (&P)[32], u64 (&N)[32], const u64 (&TQ)[4], uint32_t& wlo, uint32_t& whi)
    {
        constexpr int c = bm_class(K);
        float eLo, oLo, oHi, eHi;
        upk2(add2(P[K], TQ[c]), eLo, oLo);
        upk2(add2(P[K], TQ[c ^ 3]), eHi, oHi);
        wlo = push_sign(__fsub_rn(eLo, oLo), wlo);
        whi = push_sign(__fsub_rn(eHi, oHi), whi);
        N[K] = pk2(fmaxf(eLo, oLo), fmaxf(eHi, oHi));
        Inner<K - 1>::run(P, N, TQ, wlo, whi);
    }
};
template <>
struct Inner<-1> {
    static __device__ __forceinline__ void run(const u64 (&)[32], u64 (&)[32], const u64 (&)[4], uint32_t&, uint32_t&) {}
};

// The four branch-metric pairs TP[x] = (tab[x], tab[x ^ DELTA]) of a step, tab = {0, t1, t0, t1 + t0}, made from the pair
// T01 = (t0, t1) as it comes out of shared memory with two-wide arithmetic only (no register moves: building them with
// mov.b64 makes ptxas re-materialise a pair at nearly every use).  C01 = (0.0f, 1.0f) is read from the table blob, i.e. a
// run-time value the assembler cannot re-materialise either.  x * 0 is +-0 and m + (+-0) == m, fma(x, 1, y) == x + y rounded
// once, t0 + t1 == t1 + t0: every metric is the bit the scalar form produces.
template <int DELTA>
__device__ __forceinline__ void metric_pairs(const u64 T01, const u64 C01, u64 (&TP)[4])
{
    const u64 S01 = swap2(T01);                                         // (t1, t0)
    if constexpr (DELTA == 3) {
        const u64 T33 = add2(S01, T01);                                 // (t1 + t0, t0 + t1)
        TP[0] = mul2(T33, C01);                                         // (0, t3)
        TP[3] = swap2(TP[0]);
        TP[1] = S01;
        TP[2] = T01;
    } else if constexpr (DELTA == 1) {
        TP[0] = mul2(T01, C01);                                         // (0, t1)
        TP[1] = swap2(TP[0]);
        TP[2] = fma2(S01, C01, T01);                                    // (t1 * 0 + t0, t0 * 1 + t1) = (t0, t3)
        TP[3] = swap2(TP[2]);
    } else if constexpr (DELTA == 2) {
        const u64 C10 = swap2(C01);
        TP[2] = mul2(T01, C10);                                         // (t0, 0)
        TP[0] = swap2(TP[2]);
        TP[3] = fma2(S01, C10, T01);                                    // (t1 * 1 + t0, t0 * 0 + t1) = (t3, t1)
        TP[1] = swap2(TP[3]);
    } else {
        float t0, t1;
        upk2(T01, t0, t1);
        TP[0] = 0;                                                      // never read: adding (0, 0) is skipped
        TP[1] = pk2(t1, t1);
        TP[2] = pk2(t0, t0);
        TP[3] = add2(S01, T01);
    }
}

// one trellis step, input in layout L_Q, output in L_(Q-1) (Q = 0: L_5)
template <int Q>
__device__ __forceinline__ uint2 acs2(const u64 (&P)[32], u64 (&N)[32], const u64 T01, const u64 C01)
{
    u64 TP[4];
    uint32_t wlo = 0, whi = 0;
    if constexpr (Q == 0) {
        metric_pairs<3>(T01, C01, TP);                                  // (tab[c], tab[c ^ 3]) = (A, B) of butterfly class c
        Inner<31>::run(P, N, TP, wlo, whi);
    } else {
        metric_pairs<bm_class(1 << (Q - 1))>(T01, C01, TP);
        if constexpr (Q <= 3) {
            Groups<Q, (32 >> Q) - 1>::run(P, N, TP, wlo, whi);
        } else {
            uint32_t aLo = 0, bLo = 0, aHi = 0, bHi = 0;
            Halves<Q, 15>::run(P, N, TP, aLo, bLo, aHi, bHi);
            if constexpr (Q == 5) { wlo = (aLo << 16) | bLo; whi = (aHi << 16) | bHi; }
            else { wlo = __byte_perm(bLo, aLo, 0x5140); whi = __byte_perm(bHi, aHi, 0x5140); }
        }
    }
    return make_uint2(wlo, whi);
}

// chunk-relative soft-bit indices of step s for code rate cr; -1 = punctured (same closed forms as k_viterbi.cu)
__device__ __forceinline__ void depunc(int cr, int t, int& i0, int& i1)
{
    if (cr == C8B_CR_12) { i0 = 2 * t; i1 = 2 * t + 1; }
    else if (cr == C8B_CR_23) { int q = t >> 1, b = 3 * q; if (t & 1) { i0 = b + 2; i1 = -1; } else { i0 = b; i1 = b + 1; } }
    else if (cr == C8B_CR_34) {
        int q = t / 3, r = t - 3 * q, b = 4 * q;
        if (r == 0) { i0 = b; i1 = b + 1; } else if (r == 1) { i0 = b + 2; i1 = -1; } else { i0 = -1; i1 = b + 3; }
    } else {
        int q = t / 5, r = t - 5 * q, b = 6 * q;
        if (r == 0) { i0 = b; i1 = b + 1; }
        else if (r == 1) { i0 = b + 2; i1 = -1; }
        else if (r == 2) { i0 = -1; i1 = b + 3; }
        else if (r == 3) { i0 = b + 4; i1 = -1; }
        else { i0 = -1; i1 = b + 5; }
    }
}
__device__ __forceinline__ int used_by(int cr, int T)
{
    if (cr == C8B_CR_12) return 2 * T;
    if (cr == C8B_CR_23) return 3 * (T >> 1) + ((T & 1) ? 2 : 0);
    if (cr == C8B_CR_34) { const int q = T / 3, r = T - 3 * q; return 4 * q + (r == 0 ? 0 : r + 1); }
    const int q = T / 5, r = T - 5 * q;
    return 6 * q + (r == 0 ? 0 : r + 1);
}

// Scratch per resident CTA, in global memory:
//   surv  [C8B_DECODE_T_MAX + CS][TPB] uint2     decision words, one per trellis step and thread
//   words [(C8B_DECODE_T_MAX + 63) / 32][TPB]    decoded bits, LSB first; descrambled in place

__device__ __forceinline__ uint32_t get_byte(const uint32_t* __restrict__ words, int i)
{
    return (words[(size_t)(i >> 2) * TPB] >> (8 * (i & 3))) & 0xffu;
}

// CRC-32 (boost::crc_32_type, lib/decode_impl.h:84) over bytes [start, start+n) of the thread's word column
__device__ uint32_t crc32_words(const uint32_t* __restrict__ tab, const uint32_t* __restrict__ words, int start, int n)
{
    uint32_t c = 0xffffffffu;
    int i = start;
    const int end = start + n;
    for (; i < end && (i & 3); i++) c = tab[(c ^ get_byte(words, i)) & 0xff] ^ (c >> 8);
    for (; i + 4 <= end; i += 4) {
        const uint32_t w = words[(size_t)(i >> 2) * TPB];
        c = tab[(c ^ w) & 0xff] ^ (c >> 8);
        c = tab[(c ^ (w >> 8)) & 0xff] ^ (c >> 8);
        c = tab[(c ^ (w >> 16)) & 0xff] ^ (c >> 8);
        c = tab[(c ^ (w >> 24)) & 0xff] ^ (c >> 8);
    }
    for (; i < end; i++) c = tab[(c ^ get_byte(words, i)) & 0xff] ^ (c >> 8);
    return ~c;
}

// [fmt][len lo][len hi][MPDU][mcs] appended to the frame's PDU area by one thread
__device__ void emit_record_t(uint8_t* __restrict__ out, int& w, int cap, int& npdu, int fmt, int lenField, const uint32_t* __restrict__ words,
                              int start, int nbody, int mcs)
{
    const int rec = nbody + 4;
    if (w + rec > cap) return;
    uint8_t* o = out + w;
    o[0] = (uint8_t)fmt; o[1] = (uint8_t)(lenField & 255); o[2] = (uint8_t)(lenField >> 8);
    int i = 0;
    // bytes until the destination is word aligned, then 4 bytes per store (source words funnel-shifted into place)
    for (; i < nbody && (((uintptr_t)(o + 3 + i)) & 3); i++) o[3 + i] = (uint8_t)get_byte(words, start + i);
    for (; i + 4 <= nbody; i += 4) {
        const int k = start + i;
        const uint32_t w0 = words[(size_t)(k >> 2) * TPB], w1 = words[(size_t)((k >> 2) + 1) * TPB];
        *reinterpret_cast<uint32_t*>(o + 3 + i) = __funnelshift_r(w0, w1, 8 * (k & 3));
    }
    for (; i < nbody; i++) o[3 + i] = (uint8_t)get_byte(words, start + i);
    o[3 + nbody] = (uint8_t)mcs;
    w += rec;
    npdu++;
}

#ifndef C8B_TP_PACKED
#define C8B_TP_PACKED 0                        // 0: scalar butterflies (default, faster); 1: FADD2 forward pass (acs2) -- measured SLOWER, see below
#endif
#ifndef C8B_TP_CTAS
#define C8B_TP_CTAS 3                          // CTAs per SM (168 registers per thread, no spills; 4 would spill the metrics)
#endif
__global__ void __launch_bounds__(TPB, C8B_TP_CTAS)
k_viterbi_tp(const c8b_lut* __restrict__ lut, c8b_frame* __restrict__ frames, int nframes, const float* __restrict__ llrArena,
             int64_t nllr, uint2* __restrict__ survAll, uint32_t* __restrict__ wordsAll, size_t survPerCta, size_t wordsPerCta,
             uint8_t* __restrict__ pdu, int64_t pduStride, uint8_t* __restrict__ scram, int64_t scramStride)
{
    extern __shared__ __align__(16) uint8_t dynsm[];
    float2 (*pairs)[2][32 * ROWF2] = reinterpret_cast<float2 (*)[2][32 * ROWF2]>(dynsm);   // [warp][buffer][frame row][step]
    uint32_t* crcTab = reinterpret_cast<uint32_t*>(dynsm + sizeof(float2) * (TPB / 32) * 2 * 32 * ROWF2);
    uint32_t* relTab = crcTab + 256;                                 // [4 code rates][32]
    for (int i = threadIdx.x; i < 256; i += TPB) crcTab[i] = lut->crc32tab[i];
    for (int i = threadIdx.x; i < 128; i += TPB) {
        int a0, a1;
        depunc(i >> 5, (i & 31) < CS ? (i & 31) : 0, a0, a1);
        relTab[i] = (uint32_t)(a0 & 0xffff) | ((uint32_t)(a1 & 0xffff) << 16);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint2* __restrict__ surv = survAll + (size_t)blockIdx.x * survPerCta + threadIdx.x;        // [t * TPB]
    uint32_t* __restrict__ words = wordsAll + (size_t)blockIdx.x * wordsPerCta + threadIdx.x;  // [w * TPB]

    for (int g = blockIdx.x; g * TPB < nframes; g += gridDim.x) {
        const int f = g * TPB + threadIdx.x;
        // ---- this thread's frame ----
        int T = 0, cr = 0, total = 0, fmt = 0, len = 0, mcs = 0, ampdu = 0;
        const float* llr = llrArena;
        if (f < nframes) {
            c8b_frame* fr = frames + f;
            const int status = fr->status;
            const int t = fr->trellis;
            const int64_t loff = fr->llr_off;
            cr = fr->cr & 3; total = fr->total; fmt = fr->format; len = fr->len; mcs = fr->mcs; ampdu = fr->ampdu;
            fr->npdu = 0; fr->pdu_bytes = 0; fr->pdu_off = (int64_t)f * pduStride;
            if (status == C8B_ST_OK) {
                if (len > C8B_DECODE_B_MAX || t > C8B_DECODE_T_MAX) fr->status = C8B_ST_DECODE_RANGE;     // lib/decode_impl.cc:93-97
                else if (t > 0 && total >= 0 && loff >= 0 && loff + total <= nllr) { T = t; llr = llrArena + loff; }
            }
        }
        const int lim = min(total, used_by(cr, T));                  // soft bits this packet consumes (pad steps read 0)
        const int nraw = used_by(cr, CS);                            // soft bits per chunk of this frame
        int Tmax = T;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) Tmax = max(Tmax, __shfl_xor_sync(0xffffffffu, Tmax, o));
        if (Tmax == 0) continue;                                     // warp-uniform
        const int nch = (Tmax + CS - 1) / CS;

        // stage chunk c: every lane copies the (t0,t1) pairs of ITS frame into its shared-memory row with cp.async
        // (LDGSTS): the copies of the next chunk run under the butterflies of the current one and hold no registers.
        // relTab[cr][s] = chunk-relative soft-bit indices of step s (i0 | i1 << 16, 0xffff = punctured).  A chunk is a whole
        // number of puncture periods, so the punctured slots of a row are the same in every chunk: they are zeroed once per
        // frame and never copied; a position past the end of the packet is a zero-fill copy (src-size 0).  The table entries
        // of ten steps are fetched together (LDS.64) ahead of their twenty copies -- one dependent LDS per step stalled the
        // warp for a fifth of the kernel.
        {
            float2* r0 = pairs[warp][0] + lane * ROWF2;
            float2* r1 = pairs[warp][1] + lane * ROWF2;
#pragma unroll
            for (int i = 0; i < CS; i++) { r0[i] = make_float2(0.f, 0.f); r1[i] = make_float2(0.f, 0.f); }
            __syncwarp();
        }
        auto stage = [&](int c, int buf) {
            const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(pairs[warp][buf] + lane * ROWF2);
            const float* __restrict__ lb = llr + (size_t)c * nraw;
            const int rem = lim - c * nraw;                          // soft bits left from the start of this chunk
            const uint2* __restrict__ rt = reinterpret_cast<const uint2*>(relTab + cr * 32);
#pragma unroll
            for (int b = 0; b < CS; b += 10) {
                uint2 e[5];
#pragma unroll
                for (int k = 0; k < 5; k++) e[k] = rt[b / 2 + k];
#pragma unroll
                for (int k = 0; k < 10; k++) {
                    const uint32_t ev = (k & 1) ? e[k >> 1].y : e[k >> 1].x;
                    const int i0 = (int)(ev & 0xffffu), i1 = (int)(ev >> 16);
                    const bool in0 = i0 < rem, in1 = i1 < rem;
                    if (i0 != 0xffff)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d0 + 8 * (b + k)), "l"(lb + (in0 ? i0 : 0)), "r"(in0 ? 4 : 0));
                    if (i1 != 0xffff)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d0 + 8 * (b + k) + 4), "l"(lb + (in1 ? i1 : 0)), "r"(in1 ? 4 : 0));
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto stage_wait = [&]() { asm volatile("cp.async.wait_group 0;" ::: "memory"); __syncwarp(); };

        // ---------------- forward pass ----------------
#if C8B_TP_PACKED
        const u64 C01 = *reinterpret_cast<const u64*>(lut->pair01);    // (0.0f, 1.0f), see metric_pairs
        u64 m[32], n[32];                                              // layout L_5 at every multiple of 6 steps: pair k = states (k, k + 32)
#pragma unroll
        for (int i = 0; i < 32; i++) m[i] = pk2(-1000000000000000.0f, -1000000000000000.0f);     // lib/decode_impl.cc:171-176
        m[0] = pk2(0.0f, -1000000000000000.0f);
        stage(0, 0);
        stage_wait();
        for (int c = 0; c < nch; c++) {
            if (c + 1 < nch) stage(c + 1, (c + 1) & 1);
            const u64* __restrict__ row = reinterpret_cast<const u64*>(pairs[warp][c & 1] + lane * ROWF2);   // (t0, t1) per step
            uint2* __restrict__ sv = surv + (size_t)c * CS * TPB;
#pragma unroll 1
            for (int s = 0; s < CS; s += 6) {
                sv[(size_t)(s + 0) * TPB] = acs2<5>(m, n, row[s + 0], C01);
                sv[(size_t)(s + 1) * TPB] = acs2<4>(n, m, row[s + 1], C01);
                sv[(size_t)(s + 2) * TPB] = acs2<3>(m, n, row[s + 2], C01);
                sv[(size_t)(s + 3) * TPB] = acs2<2>(n, m, row[s + 3], C01);
                sv[(size_t)(s + 4) * TPB] = acs2<1>(m, n, row[s + 4], C01);
                sv[(size_t)(s + 5) * TPB] = acs2<0>(n, m, row[s + 5], C01);
            }
            stage_wait();
        }
#else
        float m[64], n[64];
#pragma unroll
        for (int i = 0; i < 64; i++) m[i] = -1000000000000000.0f;     // lib/decode_impl.cc:171-176
        m[0] = 0.0f;
        stage(0, 0);
        stage_wait();
        for (int c = 0; c < nch; c++) {
            if (c + 1 < nch) stage(c + 1, (c + 1) & 1);
            const float2* __restrict__ row = pairs[warp][c & 1] + lane * ROWF2;
            uint2* __restrict__ sv = surv + (size_t)c * CS * TPB;
#pragma unroll 1
            for (int s = 0; s < CS; s += 2) {
                const uint2 w0 = acs(m, n, row[s]);
                sv[(size_t)s * TPB] = w0;
                const uint2 w1 = acs(n, m, row[s + 1]);
                sv[(size_t)(s + 1) * TPB] = w1;
            }
            stage_wait();
        }
#endif

        // ---------------- traceback (lib/decode_impl.cc:282-302), final state 0 ----------------
        {
            uint32_t s = 0, acc = 0;
            constexpr int TB = 32;                                   // decision words per block; the next block is in flight
            uint2 wa[TB], wb[TB];
            auto fetch = [&](uint2 (&w)[TB], int tb) {
#pragma unroll
                for (int k = 0; k < TB; k++) w[k] = (tb >= 0 && tb + k < T) ? surv[(size_t)(tb + k) * TPB] : make_uint2(0u, 0u);
            };
            auto walk = [&](const uint2 (&w)[TB], int tb) {
#pragma unroll
                for (int k = TB - 1; k >= 0; k--) {
                    const int t = tb + k;
                    if (t < T) {
                        acc = (acc << 1) | (s >> 5);                 // decoded bit of step t = input bit of the state entered
                        const uint32_t d = ((s & 32u ? w[k].y : w[k].x) >> (s & 31u)) & 1u;
                        s = ((s & 31u) << 1) | d;
                        if ((t & 31) == 0) { words[(size_t)(t >> 5) * TPB] = acc; acc = 0; }
                    }
                }
            };
            int tb = ((Tmax - 1) / TB) * TB;
            fetch(wa, tb);
            for (; tb >= 0; tb -= 2 * TB) {
                fetch(wb, tb - TB);
                walk(wa, tb);
                fetch(wa, tb - 2 * TB);
                if (tb - TB >= 0) walk(wb, tb - TB);
            }
        }
        if (T <= 0) continue;                                        // (lanes without a frame are done; no warp-level sync below)
        const int nwords = (T + 31) >> 5;
        if (scram != nullptr) {
            uint8_t* so = scram + (size_t)f * scramStride;
            for (int i = 0; i < T && i < scramStride; i++) so[i] = (uint8_t)((words[(size_t)(i >> 5) * TPB] >> (i & 31)) & 1u);
        }

        // ---------------- descramble (lib/decode_impl.cc:304-323) ----------------
        {
            const uint32_t w0 = words[0];
            int st = 0;
#pragma unroll
            for (int i = 0; i < 7; i++) st |= (int)((w0 >> i) & 1u) << (6 - i);
            uint32_t q[6];
#pragma unroll
            for (int wq = 0; wq < 5; wq++) {
                uint32_t v = 0;
                for (int b = 0; b < 32; b++) {
                    const int fb = ((st >> 6) ^ (st >> 3)) & 1;
                    st = ((st << 1) & 0x7e) | fb;
                    v |= (uint32_t)fb << b;
                }
                q[wq] = v;
            }
            q[5] = 0;
            for (int w = 0; w < nwords; w++) {
                uint32_t v = words[(size_t)w * TPB];
                if (w == 0) v = (v ^ (q[0] << 7)) & ~0x7fu;
                else {
                    const int o = (32 * w - 7) % 127, k = o >> 5;
                    const uint32_t lo = k == 0 ? q[0] : k == 1 ? q[1] : k == 2 ? q[2] : q[3];
                    const uint32_t hi = k == 0 ? q[1] : k == 1 ? q[2] : k == 2 ? q[3] : q[4];
                    v ^= __funnelshift_r(lo, hi, o & 31);
                }
                words[(size_t)w * TPB] = v;
            }
        }

        // ---------------- packetAssemble (lib/decode_impl.cc:325-520) ----------------
        {
            uint8_t* out = pdu + (size_t)f * pduStride;
            const int cap = (int)min(pduStride, (int64_t)0x7fffffff);
            int npdu = 0, w = 0;
            if (fmt == C8B_F_VHT) {
                int procd = 16;
                if (procd < T) {
                    int bp = 2, tl = 0;                              // tl is NOT reset per subframe (:336)
                    while (true) {
                        procd += 32;
                        if (procd > T) break;
                        const int d0 = (int)get_byte(words, bp), d1 = (int)get_byte(words, bp + 1);
                        const int eof = d0 & 1;
                        tl |= ((d0 >> 2) & 1) << 12;
                        tl |= ((d0 >> 3) & 1) << 13;
                        tl |= (d0 >> 4) | (d1 << 4);
                        const int padded = (tl / 4 + ((tl % 4) != 0)) * 4;
                        procd += padded * 8;
                        if (procd > T) break;
                        bp += 4;
                        if (crc32_words(crcTab, words, bp, tl) == 558161692u) {
                            emit_record_t(out, w, cap, npdu, fmt, tl, words, bp, tl, mcs);
                            tl += 4;                                 // :415, carried into the next subframe
                        }
                        bp += padded;
                        if (eof) break;
                    }
                }
            } else if (!ampdu) {
                if (len >= 0 && 16 + 8 * len <= 32 * nwords) {
                    if (crc32_words(crcTab, words, 2, len) == 558161692u) emit_record_t(out, w, cap, npdu, fmt, len, words, 2, len, mcs);
                }
            }
            frames[f].npdu = npdu; frames[f].pdu_bytes = w;
        }
    }
}

}  // namespace

static constexpr size_t tp_smem_bytes() { return sizeof(float2) * (TPB / 32) * 2 * 32 * ROWF2 + 1024 + 512; }

// The dynamic shared-memory opt-in (> 48 KB) is a per-DEVICE function attribute: every context sets it on its own device
// at c8b_create (two contexts on two GPUs in one process, block threads created concurrently).
cudaError_t c8b_viterbi_tp_prepare(void)
{
    return cudaFuncSetAttribute(k_viterbi_tp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tp_smem_bytes());
}

size_t c8b_viterbi_tp_scratch_bytes(int num_sm)
{
    const size_t survPerCta = (size_t)(C8B_DECODE_T_MAX + CS + 2) * TPB;                 // uint2
    const size_t wordsPerCta = (size_t)((C8B_DECODE_T_MAX + 63) / 32 + 1) * TPB;        // uint32
    return (size_t)num_sm * C8B_TP_CTAS * (survPerCta * sizeof(uint2) + wordsPerCta * sizeof(uint32_t));
}

int c8b_viterbi_tp_wave(int num_sm) { return num_sm * C8B_TP_CTAS * TPB; }   // frames in one full wave

void c8b_launch_viterbi_tp(const c8b_lut* d_lut, c8b_frame* d_frames, int nframes, const float* d_llr, int64_t nllr, void* d_scratch,
                           int num_sm, uint8_t* d_pdu, int64_t pdu_stride, uint8_t* d_scram, int64_t scram_stride, cudaStream_t st)
{
    if (nframes <= 0) return;
    const size_t survPerCta = (size_t)(C8B_DECODE_T_MAX + CS + 2) * TPB;
    const size_t wordsPerCta = (size_t)((C8B_DECODE_T_MAX + 63) / 32 + 1) * TPB;
    int grid = num_sm * C8B_TP_CTAS;
    const int need = (nframes + TPB - 1) / TPB;
    if (grid > need) grid = need;
    uint2* surv = reinterpret_cast<uint2*>(d_scratch);
    uint32_t* words = reinterpret_cast<uint32_t*>(surv + (size_t)num_sm * C8B_TP_CTAS * survPerCta);
    const size_t smem = tp_smem_bytes();
    k_viterbi_tp<<<grid, TPB, smem, st>>>(d_lut, d_frames, nframes, d_llr, nllr, surv, words, survPerCta, wordsPerCta, d_pdu, pdu_stride,
                                        d_scram, scram_stride);
}
