// phy_serial.cuh -- the per-frame (once per packet) parts of the receive chain as single-thread
// routines: trigger FSM, LTF sync + CFO, L-SIG decode, format detection, HT-SIG / VHT-SIG-A / SIG-B,
// channel estimates.  One GPU thread runs them for one item / frame (k_detect.cu, k_header.cu):
// they cost ~1 % of the per-symbol and Viterbi work, so lanes are spent on frames, not inside one.
//
// The same source compiles for the host (plain C++, tests/hostsim) so the host logic is checked
// against the oracle without a GPU; the product library only ever runs it on the device.
//
// Arithmetic follows the reference operation by operation (file:line cited per routine):
//  * complex/complex division is evaluated in double and rounded once -- that is what
//    std::complex<float>::operator/ compiles to in the reference (libgcc __divsc3 on x86-64);
//  * |z| is (float)sqrt((double)re^2 + im^2) (glibc hypotf), sqrtf/float division are IEEE;
//  * atan2f: evaluated in double on the float arguments and rounded; cosf/sinf: sincosf on the device
//    (<= 2 ulp of the reference's glibc results), double-rounded in the host build;
//  * the 64-point DFT stands in for gr::fft (FFTW): double radix-2, rounded once to float.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/c80211b200.h"
#include "lut.h"

#ifdef __CUDACC__
#define C8B_HD __host__ __device__ __forceinline__
#define C8B_HDN static __host__ __device__ __noinline__
#else
#define C8B_HD inline
#define C8B_HDN static
#endif

namespace c8b {

struct cf { float re, im; };

// float ops that must not be contracted into FMAs (host build uses -ffp-contract=off)
#ifdef __CUDA_ARCH__
C8B_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
C8B_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
C8B_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
C8B_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
C8B_HD float fsqrt(float a) { return __fsqrt_rn(a); }
#else
C8B_HD float fadd(float a, float b) { return a + b; }
C8B_HD float fsub(float a, float b) { return a - b; }
C8B_HD float fmul(float a, float b) { return a * b; }
C8B_HD float fdiv(float a, float b) { return a / b; }
C8B_HD float fsqrt(float a) { return sqrtf(a); }
#endif

#ifdef __CUDA_ARCH__
C8B_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
C8B_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
C8B_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
#else
C8B_HD double dadd(double a, double b) { return a + b; }
C8B_HD double dsub(double a, double b) { return a - b; }
C8B_HD double dmul(double a, double b) { return a * b; }
#endif

C8B_HD cf mk(float r, float i) { cf z; z.re = r; z.im = i; return z; }
C8B_HD cf cadd(cf a, cf b) { return mk(fadd(a.re, b.re), fadd(a.im, b.im)); }
C8B_HD cf csub(cf a, cf b) { return mk(fsub(a.re, b.re), fsub(a.im, b.im)); }
C8B_HD cf cconj(cf a) { return mk(a.re, -a.im); }
C8B_HD cf cmul(cf a, cf b) { return mk(fsub(fmul(a.re, b.re), fmul(a.im, b.im)), fadd(fmul(a.re, b.im), fmul(a.im, b.re))); }
C8B_HD cf cscale(cf a, float s) { return mk(fmul(a.re, s), fmul(a.im, s)); }
C8B_HD cf cdivs(cf a, float s) { return mk(fdiv(a.re, s), fdiv(a.im, s)); }
C8B_HD cf cdiv(cf a, cf b)      // std::complex<float> operator/ as compiled for the reference
{
    const double ar = a.re, ai = a.im, br = b.re, bi = b.im;
    const double den = dadd(dmul(br, br), dmul(bi, bi));
    return mk((float)(dadd(dmul(ar, br), dmul(ai, bi)) / den), (float)(dsub(dmul(ai, br), dmul(ar, bi)) / den));
}
C8B_HD float cabsf_(cf a) { return (float)sqrt(dadd(dmul((double)a.re, (double)a.re), dmul((double)a.im, (double)a.im))); }
C8B_HD float cosf_(float x) { return (float)cos((double)x); }
C8B_HD float sinf_(float x) { return (float)sin((double)x); }
C8B_HD float atan2f_(float y, float x) { return (float)atan2((double)y, (double)x); }
#ifdef __CUDA_ARCH__
C8B_HD cf cis(float ph) { float sn, cs; sincosf(ph, &sn, &cs); return mk(cs, sn); }   // <= 2 ulp; the reference calls cosf / sinf
#else
C8B_HD cf cis(float ph) { return mk(cosf_(ph), sinf_(ph)); }
#endif

// ---------------------------------------------------------------------------------------------
// 64-point forward DFT, unnormalised, natural order (gr::fft::fft_complex_fwd(64) at
// lib/signal_impl.cc:121-123, lib/demod_impl.cc:541-547).  in/out may alias.
// ---------------------------------------------------------------------------------------------
C8B_HDN void fft64(const c8b_lut* L, const cf* in, cf* out)
{
    double re[64], im[64];
    for (int i = 0; i < 64; i++) {
        int r = ((i & 1) << 5) | ((i & 2) << 3) | ((i & 4) << 1) | ((i & 8) >> 1) | ((i & 16) >> 3) | ((i & 32) >> 5);
        re[r] = in[i].re; im[r] = in[i].im;
    }
    for (int len = 2; len <= 64; len <<= 1) {
        const int half = len >> 1, step = 64 / len;
        for (int b = 0; b < 64; b += len)
            for (int k = 0; k < half; k++) {
                const double wr = L->twdr[k * step], wi = L->twdi[k * step];
                const double xr = dsub(dmul(re[b + k + half], wr), dmul(im[b + k + half], wi));
                const double xi = dadd(dmul(re[b + k + half], wi), dmul(im[b + k + half], wr));
                re[b + k + half] = dsub(re[b + k], xr); im[b + k + half] = dsub(im[b + k], xi);
                re[b + k] = dadd(re[b + k], xr); im[b + k] = dadd(im[b + k], xi);
            }
    }
    for (int i = 0; i < 64; i++) out[i] = mk((float)re[i], (float)im[i]);
}

// ---------------------------------------------------------------------------------------------
// presiso at one index (examples/presiso.grc:35-229): the moving sums are evaluated with the fixed
// tree  s2[n]=v[n-1]+v[n], s4[n]=s2[n-2]+s2[n], s8, s16;  sum48[n]=(s16[n-32]+s16[n-16])+s16[n],
// sum64[n]=(s16[n-48]+s16[n-32])+(s16[n-16]+s16[n]),  v[m]=0 for m<0 (k_presiso.cu does the same
// for every sample; this routine recomputes one value on demand).
// ---------------------------------------------------------------------------------------------
C8B_HD cf conj_prod(const cf* x, int i)   // delay(16) -> multiply_conjugate_cc: x[i-16] * conj(x[i])
{
    if (i < 16) return mk(0.f, 0.f);
    const cf d = x[i - 16], c = x[i];
    return mk(fadd(fmul(d.re, c.re), fmul(d.im, c.im)), fsub(fmul(d.im, c.re), fmul(d.re, c.im)));
}
C8B_HDN cf presiso_s16(const cf* x, int n)   // tree sum of conj_prod over [n-15, n]
{
    if (n < 0) return mk(0.f, 0.f);
    cf v[16];
    for (int k = 0; k < 16; k++) v[k] = conj_prod(x, n - 15 + k);
    for (int w = 1; w < 16; w <<= 1)
        for (int k = 0; k < 16; k += 2 * w) v[k] = cadd(v[k], v[k + w]);
    return v[0];
}
C8B_HD cf presiso_conj_at(const cf* x, int i)   // moving_average_cc(48) output at i
{
    return cadd(cadd(presiso_s16(x, i - 32), presiso_s16(x, i - 16)), presiso_s16(x, i));
}

// ---------------------------------------------------------------------------------------------
// trigger (lib/trigger_impl.cc:59-117): plateau detector on preac
// ---------------------------------------------------------------------------------------------
struct TrigState { int nPlateau, fPlateau, fPlateauEnd, countDown; float conjAc; };

C8B_HD void trig_reset(TrigState& s) { s.nPlateau = 0; s.fPlateau = 0; s.fPlateauEnd = 0; s.countDown = 0; s.conjAc = 0.f; }

C8B_HD uint8_t trig_step(TrigState& s, float ac)
{
    uint8_t o = 0;
    if (ac > 0.3f) {                                              // :79
        s.nPlateau++;
        if (ac > s.conjAc) { s.conjAc = ac; o |= 0x02; }          // :82-87
        if (s.nPlateau > 20 && (s.fPlateau + s.fPlateauEnd) == 0) { s.fPlateau = 1; s.fPlateauEnd = 1; s.countDown = 80; }   // :88-93
    } else {                                                      // :95-100
        s.nPlateau = 0; s.fPlateauEnd = 0; s.conjAc = 0.0f;
    }
    if (s.fPlateau) {                                             // :101-109
        s.countDown--;
        if (s.countDown == 0) { s.fPlateau = 0; o |= 0x01; }
    }
    return o;
}

// ---------------------------------------------------------------------------------------------
// sync (lib/sync_impl.cc:92-147 SYNC state; :155-179 ltf_autoCorrelation; :181-196 ltf_cfo).
// sig = 240 samples from the trigger.  The 111 lags use the reference's running sums.
// ---------------------------------------------------------------------------------------------
struct SyncOut { int ok, mIndex; float rad, snr, rssi; };

C8B_HD float abs2(cf a) { const float m = cabsf_(a); return fmul(m, m); }   // std::abs(z) * std::abs(z)

C8B_HDN SyncOut sync_at(const cf* sig, cf conjAvg)
{
    float ac[C8B_SYNC_RES];
    cf msum = mk(0.f, 0.f);
    float s1 = 0.f, s2 = 0.f;
    for (int i = 0; i < 64; i++) {                                // :161-166
        msum = cadd(msum, cmul(sig[i], cconj(sig[i + 64])));
        s1 = fadd(s1, abs2(sig[i]));
        s2 = fadd(s2, abs2(sig[i + 64]));
    }
    float best = -1.f, bestPwr = 0.f;
    int bi = 0;
    for (int i = 0; i < C8B_SYNC_RES; i++) {                      // :167-178
        const float a = fdiv(fdiv(cabsf_(msum), fsqrt(s1)), fsqrt(s2));
        ac[i] = a;
        if (i == 0 || a > best) { best = a; bi = i; bestPwr = s1; }   // max_element: first maximum (:97)
        msum = csub(msum, cmul(sig[i], cconj(sig[i + 64])));
        s1 = fsub(s1, abs2(sig[i]));
        s2 = fsub(s2, abs2(sig[i + 64]));
        msum = cadd(msum, cmul(sig[i + 64], cconj(sig[i + 128])));
        s1 = fadd(s1, abs2(sig[i + 64]));
        s2 = fadd(s2, abs2(sig[i + 128]));
    }
    SyncOut o; o.ok = 0; o.mIndex = 0; o.rad = o.snr = o.rssi = 0.f;
    if ((double)best > 0.5) {                                     // :99
        const float thr = (float)((double)best * 0.8);            // :101 float * double -> float
        int l = bi, r = bi;
        for (int j = bi; j >= 0; j--) if (ac[j] < thr) { l = j; break; }              // :105-112
        for (int j = bi; j < C8B_SYNC_RES; j++) if (ac[j] < thr) { r = j; break; }    // :113-120
        o.ok = 1; o.mIndex = (l + r) / 2;                         // :122
        const cf* s = sig + o.mIndex;                             // ltf_cfo :181-196
        const float radStf = fdiv(atan2f_(conjAvg.im, conjAvg.re), 16.0f);
        cf csum = mk(0.f, 0.f);
        for (int i = 0; i < 64; i++) {
            const cf a = cmul(s[i], cis(fmul((float)i, radStf)));
            const cf b = cmul(s[i + 64], cis(fmul((float)(i + 64), radStf)));
            csum = cadd(csum, cmul(a, cconj(b)));
        }
        const cf c64 = cdivs(csum, 64.0f);
        const float radLtf = fdiv(atan2f_(c64.im, c64.re), 64.0f);
        o.rad = fadd(radStf, radLtf);
        const double maxD = (double)best;
        o.snr = (float)(10.0 * log10(maxD / (1.0 - maxD)));       // :126
        o.rssi = fdiv(bestPwr, 64.0f);                            // :127
    }
    return o;
}

// ---------------------------------------------------------------------------------------------
// soft Viterbi for the SIG fields, trellis <= 48 (lib/cloud80211phy.cc:2001-2088 svSigDecoder::decode).
// Same add-compare-select as the decode block; survivors kept as one decision bit per state.
// ---------------------------------------------------------------------------------------------
C8B_HDN void sig_viterbi(const c8b_lut* L, const float* llr, uint8_t* bits, int T)
{
    if (T < 0 || T > 48) return;                                  // c8p.cc:2003
    float m0[64], m1[64];
    uint64_t dec[48];
    for (int i = 0; i < 64; i++) m0[i] = -1000000000000000.0f;
    m0[0] = 0.f;
    float *pre = m0, *cur = m1;
    for (int t = 0; t < T; t++) {
        const float t0 = llr[2 * t], t1 = llr[2 * t + 1];
        float tab[4];
        tab[0] = 0.0f; tab[1] = t1; tab[2] = t0; tab[3] = fadd(t1, t0);
        uint64_t d = 0;
        for (int k = 0; k < 32; k++) {
            const int c = L->bmClass[k];
            const float a0 = fadd(pre[2 * k], tab[c]), b0 = fadd(pre[2 * k + 1], tab[c ^ 3]);       // -> state k
            const float a1 = fadd(pre[2 * k], tab[c ^ 3]), b1 = fadd(pre[2 * k + 1], tab[c]);       // -> state k+32
            float v = -1000000000000000.0f;
            if (a0 > v) v = a0;
            if (b0 > v) { v = b0; d |= 1ull << k; }
            cur[k] = v;
            v = -1000000000000000.0f;
            if (a1 > v) v = a1;
            if (b1 > v) { v = b1; d |= 1ull << (k + 32); }
            cur[k + 32] = v;
        }
        dec[t] = d;
        float* sw = pre; pre = cur; cur = sw;
    }
    int s = 0;                                                    // final state 0
    for (int t = T - 1; t >= 0; t--) {
        bits[t] = (uint8_t)(s >> 5);
        s = ((s & 31) << 1) | (int)((dec[t] >> s) & 1ull);
    }
}

// CRC-8 of HT-SIG / VHT-SIG-A (lib/cloud80211phy.cc:1367-1403 checkBitCrc8): x^8+x^2+x+1, init ones,
// output inverted, MSB first.
C8B_HD bool crc8_check(const uint8_t* bits, int len, const uint8_t* crc)
{
    unsigned c = 0xff;
    for (int i = 0; i < len; i++) {
        const unsigned top = ((c >> 7) & 1u) ^ (bits[i] & 1u);
        c = (c << 1) & 0xff;
        if (top) c ^= 0x07;
    }
    unsigned got = 0;
    for (int i = 0; i < 8; i++) got |= (unsigned)(crc[i] & 1u) << (7 - i);
    return ((~c) & 0xff) == got;
}

C8B_HD int bits_le(const uint8_t* b, int n) { int v = 0; for (int i = 0; i < n; i++) v |= ((int)b[i]) << i; return v; }

// ---------------------------------------------------------------------------------------------
// L-SIG (lib/cloud80211phy.cc:609-627 procLHSigDemodDeint, :650-728 signalCheckLegacy)
// ---------------------------------------------------------------------------------------------
C8B_HDN void lsig_demod(const c8b_lut* L, const cf* s1, const cf* s2, const cf* sig, cf* h, float* llr48)
{
    const int pb[4] = { 7, 21, 43, 57 };
    for (int q = 0; q < 4; q++) h[pb[q]] = cdivs(cadd(s1[pb[q]], s2[pb[q]]), fmul(2.0f, L->ltfL[pb[q]]));
    cf acc = cdiv(sig[7], h[7]);
    acc = csub(acc, cdiv(sig[21], h[21]));
    acc = cadd(acc, cdiv(sig[43], h[43]));
    acc = cadd(acc, cdiv(sig[57], h[57]));
    const cf ps = cconj(acc);
    const float pa = cabsf_(ps);
    for (int i = 0; i < 64; i++) {
        const int d = L->sigDemap[i];
        if (d < 0) continue;
        h[i] = cdivs(cadd(s1[i], s2[i]), fmul(2.0f, L->ltfL[i]));
        llr48[d] = cdivs(cmul(cdiv(sig[i], h[i]), ps), pa).re;
    }
}

C8B_HD int lsig_ndbps(int mcs) { const int t[8] = { 24, 36, 48, 72, 96, 144, 192, 216 }; return t[mcs & 7]; }

C8B_HD bool lsig_check(const uint8_t* b, int* mcs, int* len)
{
    if (!b[3] || b[4]) return false;                              // :652-660
    int par = 0;
    for (int i = 0; i < 17; i++) par += b[i];
    if ((par & 1) != (int)b[17]) return false;
    const int rmap[8] = { 6, 4, 2, 0, 7, 5, 3, 1 };               // R1-R3 (R4 = 1) -> rate index
    *mcs = rmap[b[0] | (b[1] << 1) | (b[2] << 2)];
    *len = bits_le(b + 5, 12);
    return *len >= 14 && *len <= 4095;
}

// signal block, S_DEMOD (lib/signal_impl.cc:108-162): in = samples from the sync index (>= 224)
C8B_HDN int signal_at(const c8b_lut* L, const cf* in, float rad, cf* h, int* mcs, int* len, int* nsamp)
{
    cf f1[64], f2[64], fs[64];
    for (int i = 0; i < 64; i++) {                                // :115-120
        f1[i] = cmul(in[C8B_SYM_SHIFT + i], cis(fmul((float)(i + C8B_SYM_SHIFT), rad)));
        f2[i] = cmul(in[C8B_SYM_SHIFT + 64 + i], cis(fmul((float)(i + C8B_SYM_SHIFT + 64), rad)));
        fs[i] = cmul(in[C8B_SYM_SHIFT + 144 + i], cis(fmul((float)(i + C8B_SYM_SHIFT + 144), rad)));
    }
    fft64(L, f1, f1); fft64(L, f2, f2); fft64(L, fs, fs);
    for (int i = 0; i < 64; i++) h[i] = mk(0.f, 0.f);
    float llr[48];
    uint8_t bits[24];
    lsig_demod(L, f1, f2, fs, h, llr);
    sig_viterbi(L, llr, bits, 24);
    if (!lsig_check(bits, mcs, len)) return 0;
    const int ndbps = lsig_ndbps(*mcs);
    *nsamp = ((*len * 8 + 22 + ndbps - 1) / ndbps) * 80;          // :128
    return 1;
}

// ---------------------------------------------------------------------------------------------
// detection of the first frame of one item: trigger FSM -> sync (IDLE/SYNC states,
// lib/sync_impl.cc:73-147) -> signal (S_TRIGGER/S_DEMOD, lib/signal_impl.cc:75-162), evaluated over
// the item from reset state.  preac = presiso output for the item.  Fills the detection fields of f.
// ---------------------------------------------------------------------------------------------
C8B_HD void frame_clear(c8b_frame* f, int item, int status)
{
    f->status = status; f->item = item;
    f->trig_idx = 0; f->sync_idx = 0; f->rad = f->snr = f->rssi = f->cfo_hz = 0.f;
    f->l_mcs = f->l_len = f->nsamp = 0;
    f->format = f->mcs = f->len = f->cr = f->ampdu = f->nss = f->nsym = f->nsymsamp = f->ncbps = f->ndbps = 0;
    f->trellis = f->total = f->data_off = 0; f->sssnr0 = f->sssnr1 = 0.f;
    f->llr_off = 0; f->pdu_off = 0; f->npdu = f->pdu_bytes = 0;
}

// f[0..maxf): frame records of this item, h[0..maxf*64): their legacy channels.  Frames are found in stream order
// exactly as the blocks would (S_COPY swallows the sync flags inside a copied frame, lib/signal_impl.cc:164-192);
// unused records get C8B_ST_EMPTY, an item without any accepted frame reports its drop code in record 0.
// mask (may be null): bit (i & 31) of mask[i >> 5] = (preac[i] > 0.3f), written by k_presiso; lets the scan jump over
// 32 samples at a time while the trigger is idle (no plateau, no count-down): such a stretch leaves the FSM in its
// reset state, so skipping it is exact.
C8B_HDN void detect_item(const c8b_lut* L, const cf* x, const float* preac, int n, int item, int maxf, c8b_frame* f, cf* h,
                         const uint32_t* mask = nullptr, c8b_scan* sc = nullptr)
{
    TrigState ts;
    trig_reset(ts);
    const bool live = sc && !sc->flush;                           // stop at the first event that needs more samples
    const int from = sc ? sc->from : 0;
    int latch = -1, skipUntil = 0, nTrig = 0, nEv = 0, nLsigFail = 0, pos = sc ? sc->pos0 : 0, nf = 0;
    int safe = from, posS = pos, nfS = 0, stalled = 0;
    bool syncStalled = false, sigStalled = false, done = false;
    for (int k = 0; k < maxf; k++) frame_clear(f + k, item, C8B_ST_EMPTY);
    for (int i = from; i < n && !done; i++) {
        if (mask && (i & 31) == 0 && i + 32 <= n && ts.fPlateau == 0 && mask[i >> 5] == 0u) {
            ts.nPlateau = 0; ts.fPlateauEnd = 0; ts.conjAc = 0.0f;   // what 32 sub-threshold samples leave behind
            i += 31;
            if (i + 1 >= skipUntil) { safe = i + 1; posS = pos; nfS = nf; }
            continue;
        }
        if (ts.nPlateau == 0 && ts.fPlateau == 0 && i >= skipUntil) { safe = i; posS = pos; nfS = nf; }
        const uint8_t fl = trig_step(ts, preac[i]);
        if (fl == 0 || i < skipUntil || syncStalled) continue;
        if (fl & 0x01) {
            nTrig++;
            if (n - i < C8B_SYNC_BUF) {                           // sync_impl.cc:94 not satisfied yet / ever
                if (live) { stalled = 1; done = true; } else syncStalled = true;
                continue;
            }
            const cf cj = latch >= 0 ? presiso_conj_at(x, latch) : mk(0.f, 0.f);
            const SyncOut so = sync_at(x + i, cj);
            skipUntil = i + C8B_SYNC_RES;
            if (!so.ok) continue;
            nEv++;
            const int idx = i + so.mIndex;
            if (sigStalled || idx < pos) continue;                // swallowed by S_COPY / skipped 80
            if (n - idx < 224) {                                  // signal_impl.cc:110 stall
                if (live) { stalled = 1; done = true; } else sigStalled = true;
                continue;
            }
            int mcs = 0, len = 0, nsamp = 0;
            cf* hk = h + nf * 64;
            if (!signal_at(L, x + idx, so.rad, hk, &mcs, &len, &nsamp)) { nLsigFail++; pos = idx + 80; continue; }
            c8b_frame* fk = f + nf;
            nf++;
            fk->trig_idx = i; fk->sync_idx = idx; fk->rad = so.rad; fk->snr = so.snr; fk->rssi = so.rssi;
            fk->cfo_hz = fmul(so.rad, 3183098.8618379068f);       // signal_impl.cc:135
            fk->l_mcs = mcs; fk->l_len = len; fk->nsamp = nsamp;
            pos = idx + 224 + nsamp;
            if (pos > n) { fk->status = C8B_ST_TRUNC; done = true; stalled = 1; }   // S_COPY can never finish: nothing after it
            else fk->status = C8B_ST_OK;
            if (nf >= maxf) { done = true; if (!stalled) stalled = 2; }
        } else if (fl & 0x02) {
            latch = i;
        }
    }
    if (!done && ts.nPlateau == 0 && ts.fPlateau == 0 && n >= skipUntil) { safe = n; posS = pos; nfS = nf; }
    if (sc) { sc->safe = safe; sc->pos = posS; sc->nf = nfS; sc->stalled = stalled; }
    if (nf == 0) f->status = nTrig == 0 ? C8B_ST_NO_TRIGGER : nEv == 0 ? C8B_ST_SYNC : (nLsigFail ? C8B_ST_LSIG : C8B_ST_TRUNC);
}

// ---------------------------------------------------------------------------------------------
// demod header states (lib/demod_impl.cc:72-219): format detection, SIG parsers, channel estimate
// ---------------------------------------------------------------------------------------------
struct Mod { int format, sumu, ampdu, nSym, nSymSamp, nSD, nSP, nSS, nLTF, mcs, len, mod, cr, nBPSCS, nDBPS, nCBPS, nCBPSS; };

C8B_HD void rate_fields(Mod* m)                                   // tail of modParserHt/Vht (c8p.cc:1054-1088, 1282-1323)
{
    m->nSD = 52; m->nSP = 4;
    m->nCBPSS = m->nBPSCS * m->nSD;
    m->nCBPS = m->nCBPSS * m->nSS;
    switch (m->cr) {
    case C8B_CR_12: m->nDBPS = m->nCBPS / 2; break;
    case C8B_CR_23: m->nDBPS = (m->nCBPS * 2) / 3; break;
    case C8B_CR_34: m->nDBPS = (m->nCBPS * 3) / 4; break;
    case C8B_CR_56: m->nDBPS = (m->nCBPS * 5) / 6; break;
    default: break;
    }
    if (m->nSS == 1) m->nLTF = 1; else if (m->nSS == 2) m->nLTF = 2; else if (m->nSS == 3 || m->nSS == 4) m->nLTF = 4;
}

C8B_HD int nsym_of(int len, int extra, int ndbps) { const int b = len * 8 + extra; return b / ndbps + ((b % ndbps) != 0); }

C8B_HD void parse_l(int mcs, int len, Mod* m)                     // signalParserL c8p.cc:773-850
{
    const int md[8] = { 0, 0, 2, 2, 3, 3, 4, 4 };
    const int cr[8] = { C8B_CR_12, C8B_CR_34, C8B_CR_12, C8B_CR_34, C8B_CR_12, C8B_CR_34, C8B_CR_23, C8B_CR_34 };
    const int nb[8] = { 1, 1, 2, 2, 4, 4, 6, 6 };
    m->mcs = mcs;
    if (mcs >= 0 && mcs < 8) { m->mod = md[mcs]; m->cr = cr[mcs]; m->nBPSCS = nb[mcs]; m->nDBPS = lsig_ndbps(mcs); m->nCBPS = 48 * nb[mcs]; }
    m->len = len; m->nCBPSS = m->nCBPS; m->nSD = 48; m->nSP = 4; m->nSS = 1; m->sumu = 0; m->nLTF = 0;
    m->format = C8B_F_L; m->nSymSamp = 80; m->ampdu = 0;
    m->nSym = nsym_of(len, 22, m->nDBPS);
}

C8B_HD bool check_ht(const uint8_t* b)                            // signalCheckHt c8p.cc:730-751
{
    if (b[26] != 1) return false;
    if (!crc8_check(b, 34, b + 34)) return false;
    return (b[5] + b[6] + b[7] + b[28] + b[29] + b[30] + b[32] + b[33]) == 0;
}
C8B_HD bool check_vhta(const uint8_t* b)                          // signalCheckVhtA c8p.cc:753-771
{
    if (b[2] != 1 || b[23] != 1 || b[33] != 1) return false;
    if (!crc8_check(b, 34, b + 34)) return false;
    return (b[0] + b[1]) == 0;
}

C8B_HD void parse_ht(const uint8_t* b, Mod* m)                    // signalParserHt c8p.cc:852-998
{
    const int md[8] = { 0, 2, 2, 3, 3, 4, 4, 4 }, nb[8] = { 1, 2, 2, 4, 4, 6, 6, 6 };
    const int cr[8] = { C8B_CR_12, C8B_CR_12, C8B_CR_34, C8B_CR_12, C8B_CR_34, C8B_CR_23, C8B_CR_34, C8B_CR_56 };
    const int mcs = bits_le(b, 7), len = bits_le(b + 8, 16);
    m->format = C8B_F_HT; m->sumu = 0;
    m->nSymSamp = b[31] ? 72 : 80;
    m->ampdu = b[27] ? 1 : 0;
    m->mcs = mcs;
    m->mod = md[mcs & 7]; m->nBPSCS = nb[mcs & 7]; m->cr = cr[mcs & 7];
    m->len = len;
    m->nSS = mcs / 8 + 1;
    rate_fields(m);
    m->nSym = nsym_of(len, 22, m->nDBPS);
}

C8B_HD void mod_vht(int mcs, Mod* m)                              // modParserVht c8p.cc:1225-1323
{
    const int md[10] = { 0, 2, 2, 3, 3, 4, 4, 4, 5, 5 }, nb[10] = { 1, 2, 2, 4, 4, 6, 6, 6, 8, 8 };
    const int cr[10] = { C8B_CR_12, C8B_CR_12, C8B_CR_34, C8B_CR_12, C8B_CR_34, C8B_CR_23, C8B_CR_34, C8B_CR_56, C8B_CR_34, C8B_CR_56 };
    m->mcs = mcs;
    if (mcs >= 0 && mcs < 10) { m->mod = md[mcs]; m->nBPSCS = nb[mcs]; m->cr = cr[mcs]; }
    rate_fields(m);
}

C8B_HD void parse_vhta(const uint8_t* b, Mod* m)                  // signalParserVhtA c8p.cc:1090-1178
{
    const int gid = bits_le(b + 4, 6);
    m->format = C8B_F_VHT;
    m->nSymSamp = b[24] ? 72 : 80;
    m->ampdu = 1;
    if (gid == 0 || gid == 63) {
        m->sumu = 0;
        m->nSS = bits_le(b + 10, 3) + 1;
        mod_vht(bits_le(b + 28, 4), m);
    } else {
        m->sumu = 1; m->nLTF = 2; m->nSS = 1; m->nSD = 52; m->nSP = 4;
    }
}

C8B_HD void parse_vhtb(const uint8_t* b, Mod* m)                  // signalParserVhtB c8p.cc:1180-1223
{
    // An MCS field outside 0..9 leaves nDBPS unset (c8p.cc:1225-1294 default case; the reference then divides by a stale
    // member or by zero): such a frame is dropped at the sanity check (len = nSym = -1), as the oracle does.
    if (m->sumu) {
        const int len = bits_le(b, 16), mcs = bits_le(b + 16, 4);
        mod_vht(mcs, m);
        m->nLTF = 2;
        if (m->nDBPS <= 0) { m->len = -1; m->nSym = -1; return; }
        m->len = len * 4;
        m->nSym = nsym_of(m->len, 22, m->nDBPS);
    } else if ((b[17] + b[18] + b[19]) == 3) {
        if (m->nDBPS <= 0) { m->len = -1; m->nSym = -1; return; }
        m->len = bits_le(b, 17) * 4;
        m->nSym = nsym_of(m->len, 22, m->nDBPS);
    } else {
        if ((unsigned)bits_le(b, 20) == 0b01000010001011100000u) { m->len = 0; m->nSym = 0; }   // NDP
        else { m->len = -1; m->nSym = -1; }
    }
}

C8B_HD bool nl_null(int i) { return i == 0 || (i >= 29 && i <= 35); }
C8B_HD bool l_null(int i) { return i == 0 || (i >= 27 && i <= 37); }
C8B_HD bool is_pilot(int i) { return i == 7 || i == 21 || i == 43 || i == 57; }

// HT-SIG / VHT-SIG-A demod (lib/cloud80211phy.cc:629-648 procNLSigDemodDeint)
C8B_HDN void nlsig_demod(const c8b_lut* L, const cf* s1, const cf* s2, const cf* h, float* llrht, float* llrvht)
{
    cf a1 = cdiv(s1[7], h[7]), a2 = cdiv(s2[7], h[7]);
    a1 = csub(a1, cdiv(s1[21], h[21])); a2 = csub(a2, cdiv(s2[21], h[21]));
    a1 = cadd(a1, cdiv(s1[43], h[43])); a2 = cadd(a2, cdiv(s2[43], h[43]));
    a1 = cadd(a1, cdiv(s1[57], h[57])); a2 = cadd(a2, cdiv(s2[57], h[57]));
    const cf p1 = cconj(a1), p2 = cconj(a2);
    const float m1 = cabsf_(p1), m2 = cabsf_(p2);
    for (int i = 0; i < 64; i++) {
        const int d = L->sigDemap[i];
        if (d < 0) continue;
        const cf q1 = cdivs(cmul(cdiv(s1[i], h[i]), p1), m1);
        const cf q2 = cdivs(cmul(cdiv(s2[i], h[i]), p2), m2);
        llrht[d] = q1.im; llrht[d + 48] = q2.im;
        llrvht[d] = q1.re; llrvht[d + 48] = q2.im;
    }
}

// BCC encoder (lib/cloud80211phy.cc:2622-2644), used by the VHT-SIG-B SNR estimate
C8B_HD void bcc_encode(const uint8_t* in, uint8_t* out, int len)
{
    int st = 0;
    for (int i = 0; i < len; i++) {
        st = ((st << 1) & 0x7e) | in[i];
        int a = st & 0155, b = st & 0117, pa = 0, pb = 0;
        for (int q = 0; q < 7; q++) { pa ^= (a >> q) & 1; pb ^= (b >> q) & 1; }
        out[2 * i] = (uint8_t)pa; out[2 * i + 1] = (uint8_t)pb;
    }
}

// Everything demod does before the per-symbol loop.  rot(k) must return the CFO-compensated sample k
// of the signal block's output stream (k = 0 is frame sample 400).  nvalid = nsamp + 320 (S_PAD).
// hinv[64]: reciprocal of the channel the data symbols are equalised with (0 on unused bins).
// Returns the frame status; fills format .. sssnr of f.
template <class Rot>
C8B_HDN int demod_header(const c8b_lut* L, Rot rot, int nsamp, int lmcs, int llen, const cf* hl, int mupos, c8b_frame* f, cf* hinv)
{
    Mod m;
    m.format = m.sumu = m.ampdu = m.nSym = m.nSymSamp = m.nSD = m.nSP = m.nSS = m.nLTF = 0;
    m.mcs = m.len = m.mod = m.cr = m.nBPSCS = m.nDBPS = m.nCBPS = m.nCBPSS = 0;
    cf HNL[64], fo1[64], fo2[64];
    for (int i = 0; i < 64; i++) HNL[i] = mk(0.f, 0.f);
    const int nsig = nsamp + 320;
    int pos = 0, trellis = 0;
    float sssnr = 0.f;
    bool legacy = lmcs > 0;                                       // demod_impl.cc:93-100
    auto win = [&](int start, cf* out) { for (int i = 0; i < 64; i++) out[i] = rot(start + C8B_SYM_SHIFT + i); fft64(L, out, out); };
    if (!legacy) {                                                // DEMOD_S_FORMAT :106-148
        if (nsig < 160) return C8B_ST_TRUNC;
        uint8_t vb[48], hb[48];
        float llrht[96], llrvht[96];
        win(0, fo1); win(80, fo2);
        nlsig_demod(L, fo1, fo2, hl, llrht, llrvht);
        sig_viterbi(L, llrvht, vb, 48);
        if (check_vhta(vb)) {                                     // DEMOD_S_VHT :150-178
            parse_vhta(vb, &m);
            pos = 160;
            const int need = 80 + m.nLTF * 80 + 80;
            if (nsig - pos < need) return C8B_ST_TRUNC;
            // nonLegacyChanEstimate(&inSig1[80]) :344-411
            if (m.sumu) {
                win(pos + 80, fo1); win(pos + 160, fo2);
                for (int i = 0; i < 64; i++) {
                    if (nl_null(i)) continue;
                    if (mupos == 0) HNL[i] = cdivs(csub(fo1[i], fo2[i]), fmul(L->ltfNL[i], 2.0f));
                    else HNL[i] = cdivs(cadd(cdivs(fo1[i], L->ltfNL[i]), cdivs(fo2[i], L->ltfNL22[i])), 2.0f);
                }
            } else {
                win(pos + 80, fo1);
                for (int i = 0; i < 64; i++) if (!nl_null(i)) HNL[i] = cdivs(fo1[i], L->ltfNL[i]);
            }
            // vhtSigBDemod(&inSig1[80 + nLTF*80]) :449-505
            {
                cf sig1[64], bq[52];
                float inted[52], coded[52];
                uint8_t sb[26], enc[52];
                win(pos + 80 + m.nLTF * 80, fo1);
                for (int i = 0; i < 64; i++) if (!nl_null(i)) sig1[i] = cdiv(fo1[i], HNL[i]);
                const cf ps = cconj(cadd(cadd(csub(sig1[7], sig1[21]), sig1[43]), sig1[57]));
                const float pa = cabsf_(ps);
                for (int i = 0; i < 64; i++) {
                    const int d = L->binToDataNL[i];
                    if (d == 255) continue;
                    bq[d] = cdivs(cmul(sig1[i], ps), pa);
                    inted[d] = bq[d].re;
                }
                for (int i = 0; i < 52; i++) coded[L->deintNL[0][0][i]] = inted[i];     // mapDeintVhtSigB20
                sig_viterbi(L, coded, sb, 26);
                bcc_encode(sb, enc, 26);
                double np = 0.0;
                for (int i = 0; i < 52; i++) {                    // procIntelVhtB20 + noise power :488-504
                    const uint8_t bit = enc[L->deintNL[0][0][i]];
                    const cf e = mk(fsub(bq[i].re, bit ? 1.0f : -1.0f), bq[i].im);
                    np += (double)fadd(fmul(e.re, e.re), fmul(e.im, e.im));
                }
                sssnr = (float)(log10(52.0 / np) * 10.0);
                parse_vhtb(sb, &m);
            }
            const int nl = (llen * 8 + 22 + 23) / 24;
            const bool ok = m.len >= 0 && m.len <= 4095 && m.nSS <= 2 && (nl * 80) >= (m.nSym * m.nSymSamp + 160 + 80 + m.nLTF * 80 + 80);
            pos += need;
            if (!ok) return C8B_ST_FORMAT;
            trellis = m.nSym * m.nDBPS;
        } else {
            sig_viterbi(L, llrht, hb, 48);
            if (check_ht(hb)) {                                   // DEMOD_S_HT :180-205
                parse_ht(hb, &m);
                pos = 160;
                const int need = 80 + m.nLTF * 80;
                if (nsig - pos < need) return C8B_ST_TRUNC;
                win(pos + 80, fo1);
                for (int i = 0; i < 64; i++) if (!nl_null(i)) HNL[i] = cdivs(fo1[i], L->ltfNL[i]);
                const int nl = (llen * 8 + 22 + 23) / 24;
                const bool ok = m.len > 0 && m.len <= 4095 && m.nSS <= 2 && (nl * 80) >= (m.nSym * m.nSymSamp + 160 + 80 + m.nLTF * 80);
                pos += need;
                if (!ok) return C8B_ST_FORMAT;
                trellis = m.len * 8 + 22;
            } else {
                legacy = true;
            }
        }
    }
    if (legacy) {                                                 // DEMOD_S_LEGACY :207-219
        parse_l(lmcs, llen, &m);
        trellis = m.len * 8 + 22;
    }
    // DEMOD_S_WRTAG :221-277
    f->format = m.format; f->mcs = m.mcs; f->len = m.len; f->cr = m.cr; f->ampdu = m.ampdu;
    f->nss = m.nSS; f->nsym = m.nSym; f->nsymsamp = m.nSymSamp; f->ncbps = m.nCBPS; f->ndbps = m.nDBPS;
    f->trellis = trellis; f->total = m.nSym * m.nCBPS; f->data_off = pos;
    f->sssnr0 = (m.format == C8B_F_VHT) ? sssnr : 0.f; f->sssnr1 = 0.f;
    for (int i = 0; i < 64; i++) {
        const cf hh = (m.format == C8B_F_L) ? hl[i] : HNL[i];
        const bool used = (m.format == C8B_F_L) ? !l_null(i) : !nl_null(i);
        if (used) { const double den = dadd(dmul((double)hh.re, (double)hh.re), dmul((double)hh.im, (double)hh.im)); hinv[i] = mk((float)((double)hh.re / den), (float)(-(double)hh.im / den)); }
        else hinv[i] = mk(0.f, 0.f);
    }
    if (m.nSym == 0) { f->total = 1024; return C8B_ST_NDP; }
    if (m.nSS != 1) return C8B_ST_FORMAT;                         // 2-stream frames need the demod2 path
    // DEMOD_S_DEMOD needs every symbol inside the copied stream ("(o1 + nSymSamp) < d_nProc", :283)
    if (pos + m.nSym * m.nSymSamp > nsig) return C8B_ST_TRUNC;
    return C8B_ST_OK;
}

// ---------------------------------------------------------------------------------------------
// demod2 header states (lib/demod2_impl.cc:72-277, :350-469, :632-758): as demod_header, for the
// 2-antenna block.  rot0/rot1 = CFO-compensated streams of antenna 0/1 (signal2 copies both,
// lib/signal2_impl.cc:164-192); format detection uses antenna 0 only.  Outputs, besides the frame
// fields: hinv[64] when the frame is 1-stream (equalised as in the SISO block) and, for 2 streams,
//   w2[4*bin + 2a + r] : s_a = F_ant0 * w2[..2a] + F_ant1 * w2[..2a+1]   ((H^H H)^-1 H^H folded, :498-501)
//   w2[256 + q], w2[260 + q] : conj pilot references of stream 0 / 1 from the first LTF (:432-463),
//                              q = slot in the d_pilot[] order {43, 57, 7, 21}.
// ---------------------------------------------------------------------------------------------
struct cd { double re, im; };
C8B_HD cd dmk(double r, double i) { cd z; z.re = r; z.im = i; return z; }
C8B_HD cd dcmul(cd a, cd b) { return dmk(dsub(dmul(a.re, b.re), dmul(a.im, b.im)), dadd(dmul(a.re, b.im), dmul(a.im, b.re))); }
C8B_HD cd dcadd(cd a, cd b) { return dmk(dadd(a.re, b.re), dadd(a.im, b.im)); }
C8B_HD cd dconj(cd a) { return dmk(a.re, -a.im); }
C8B_HD cd tod(cf a) { return dmk((double)a.re, (double)a.im); }
C8B_HD cf tof(cd a) { return mk((float)a.re, (float)a.im); }

// Per-bin noise power for the MMSE option of the 2x2 equaliser (cfg.mmse): sync's tags give the mean power per time sample of
// signal + noise (rssi, lib/sync_impl.cc:127) and their ratio (snr in dB, :126), so the noise of one bin of the unnormalised
// 64-point DFT is 64 rssi / (1 + 10^(snr/10)).  0 (also for a noiseless capture, where the reference's snr is NaN) = the
// reference's zero-forcing arithmetic, untouched.
C8B_HD float mmse_sigma2(int mmse, float snr_db, float rssi)
{
    if (!mmse || !(snr_db == snr_db) || !(rssi > 0.f)) return 0.f;
    return (float)(64.0 * (double)rssi / (1.0 + pow(10.0, (double)snr_db / 10.0)));
}
// (H^H H + sigma2 I)^-1 in place of the reference's (H^H H)^-1 (lib/demod2_impl.cc:410-429), rows rescaled so that each
// stream keeps unit gain on itself (unbiased MMSE: the soft demapper's decision regions stay where they are).  a, b, c, d =
// the Gram matrix as the reference forms it; hi[0..3] = the inverse in the reference's layout.
C8B_HD void gram_inverse(cf a, cf b, cf c, cf d, float sigma2, cf* hi)
{
    cf ar = a, dr = d;
    if (sigma2 > 0.f) { ar.re = fadd(a.re, sigma2); dr.re = fadd(d.re, sigma2); }
    const cf inv = cdiv(mk(1.0f, 0.0f), csub(cmul(ar, dr), cmul(b, c)));
    hi[0] = cmul(inv, dr); hi[1] = cmul(mk(-inv.re, -inv.im), b);
    hi[2] = cmul(mk(-inv.re, -inv.im), c); hi[3] = cmul(inv, ar);
    if (sigma2 > 0.f) {
        // [s1 s2] = [t1 t2] hi with [t1 t2] = [x1 x2] [[a, b], [c, d]]: gain of stream k on itself = (M hi)_kk
        const float g0 = fadd(cmul(a, hi[0]).re, cmul(b, hi[2]).re), g1 = fadd(cmul(c, hi[1]).re, cmul(d, hi[3]).re);
        if (g0 > 0.f) { hi[0] = cdivs(hi[0], g0); hi[2] = cdivs(hi[2], g0); }
        if (g1 > 0.f) { hi[1] = cdivs(hi[1], g1); hi[3] = cdivs(hi[3], g1); }
    }
}

template <class Rot>
C8B_HDN int demod_header2(const c8b_lut* L, Rot rot0, Rot rot1, int nsamp, int lmcs, int llen, const cf* hl, c8b_frame* f, cf* hinv, cf* w2,
                          float sigma2 = 0.f)
{
    Mod m;
    m.format = m.sumu = m.ampdu = m.nSym = m.nSymSamp = m.nSD = m.nSP = m.nSS = m.nLTF = 0;
    m.mcs = m.len = m.mod = m.cr = m.nBPSCS = m.nDBPS = m.nCBPS = m.nCBPSS = 0;
    cf H[64][4], HI[64][4], f1[64], f2[64], f12[64], f22[64], pnl[4], pnl2[4];
    for (int i = 0; i < 64; i++) for (int k = 0; k < 4; k++) { H[i][k] = mk(0.f, 0.f); HI[i][k] = mk(0.f, 0.f); }
    for (int q = 0; q < 4; q++) { pnl[q] = mk(0.f, 0.f); pnl2[q] = mk(0.f, 0.f); }
    const int nsig = nsamp + 320;
    int pos = 0, trellis = 0;
    float sssnr0 = 0.f, sssnr1 = 0.f;
    bool legacy = lmcs > 0;
    auto win0 = [&](int start, cf* out) { for (int i = 0; i < 64; i++) out[i] = rot0(start + C8B_SYM_SHIFT + i); fft64(L, out, out); };
    auto win1 = [&](int start, cf* out) { for (int i = 0; i < 64; i++) out[i] = rot1(start + C8B_SYM_SHIFT + i); fft64(L, out, out); };
    // zero-forcing combine of one bin, the reference's operation order (:498-501)
    auto zf = [&](int i, cf a1, cf a2, cf& s1, cf& s2) {
        const cf t1 = cadd(cmul(a1, cconj(H[i][0])), cmul(a2, cconj(H[i][1])));
        const cf t2 = cadd(cmul(a1, cconj(H[i][2])), cmul(a2, cconj(H[i][3])));
        s1 = cadd(cmul(t1, HI[i][0]), cmul(t2, HI[i][2]));
        s2 = cadd(cmul(t1, HI[i][1]), cmul(t2, HI[i][3]));
    };
    auto chan_estimate = [&](int start) {                          // nonLegacyChanEstimate :350-469
        if (m.nSS == 1) {
            if (m.nLTF == 1) { win0(start, f1); for (int i = 0; i < 64; i++) if (!nl_null(i)) H[i][0] = cdivs(f1[i], L->ltfNL[i]); }
        } else if (m.nSS == 2) {
            win0(start, f1); win1(start, f2); win0(start + 80, f12); win1(start + 80, f22);
            for (int i = 0; i < 64; i++) {
                if (nl_null(i)) continue;
                const float l2 = fmul(L->ltfNL[i], 0.5f);             // LTF_NL_28_F_FLOAT2
                H[i][0] = cscale(csub(f1[i], f12[i]), l2); H[i][1] = cscale(csub(f2[i], f22[i]), l2);
                H[i][2] = cscale(cadd(f1[i], f12[i]), l2); H[i][3] = cscale(cadd(f2[i], f22[i]), l2);
            }
            const int pb[4] = { 7, 21, 43, 57 }, slot[4] = { 2, 3, 0, 1 };
            if (m.format == C8B_F_VHT)                                // pilot tones interpolated :391-409
                for (int q = 0; q < 4; q++) for (int k = 0; k < 4; k++) H[pb[q]][k] = cdivs(cadd(H[pb[q] - 1][k], H[pb[q] + 1][k]), 2.0f);
            for (int i = 0; i < 64; i++) {
                if (nl_null(i)) continue;
                const cf a = cadd(cmul(H[i][0], cconj(H[i][0])), cmul(H[i][1], cconj(H[i][1])));
                const cf b = cadd(cmul(H[i][0], cconj(H[i][2])), cmul(H[i][1], cconj(H[i][3])));
                const cf c = cadd(cmul(H[i][2], cconj(H[i][0])), cmul(H[i][3], cconj(H[i][1])));
                const cf d = cadd(cmul(H[i][2], cconj(H[i][2])), cmul(H[i][3], cconj(H[i][3])));
                gram_inverse(a, b, c, d, sigma2, HI[i]);
            }
            for (int q = 0; q < 4; q++) {
                cf t1, t2;
                zf(pb[q], f1[pb[q]], f2[pb[q]], t1, t2);
                if (q == 3) { t1 = mk(-t1.re, -t1.im); t2 = mk(-t2.re, -t2.im); }
                pnl[slot[q]] = cconj(t1); pnl2[slot[q]] = cconj(t2);
            }
        }
    };
    if (!legacy) {                                                // DEMOD_S_FORMAT :104-147
        if (nsig < 160) return C8B_ST_TRUNC;
        uint8_t vb[48], hb[48];
        float llrht[96], llrvht[96];
        win0(0, f1); win0(80, f2);
        nlsig_demod(L, f1, f2, hl, llrht, llrvht);
        sig_viterbi(L, llrvht, vb, 48);
        if (check_vhta(vb)) {                                     // DEMOD_S_VHT :149-178
            parse_vhta(vb, &m);
            pos = 160;
            const int need = 80 + m.nLTF * 80 + 80;
            if (nsig - pos < need) return C8B_ST_TRUNC;
            chan_estimate(pos + 80);
            {                                                     // vhtSigBDemod :632-758
                cf s1[64], s2[64], q0[52], q1[52];
                float inted[52], coded[52];
                uint8_t sb[26], enc[52];
                const int st = pos + 80 + m.nLTF * 80;
                bool have = true;
                if (m.nSS == 1) {
                    win0(st, f1);
                    for (int i = 0; i < 64; i++) if (!nl_null(i)) s1[i] = cdiv(f1[i], H[i][0]);
                    const cf ps = cconj(cadd(cadd(csub(s1[7], s1[21]), s1[43]), s1[57]));
                    const float pa = cabsf_(ps);
                    for (int i = 0; i < 64; i++) { const int d = L->binToDataNL[i]; if (d == 255) continue; q0[d] = cdivs(cmul(s1[i], ps), pa); inted[d] = q0[d].re; }
                } else if (m.nSS == 2) {
                    win0(st, f1); win1(st, f2);
                    for (int i = 0; i < 64; i++) if (!nl_null(i)) zf(i, f1[i], f2[i], s1[i], s2[i]);
                    cf acc = cmul(s1[7], pnl[2]);
                    acc = csub(acc, cmul(s1[21], pnl[3])); acc = cadd(acc, cmul(s1[43], pnl[0])); acc = cadd(acc, cmul(s1[57], pnl[1]));
                    acc = cadd(acc, cmul(s2[7], pnl2[2])); acc = csub(acc, cmul(s2[21], pnl2[3]));
                    acc = cadd(acc, cmul(s2[43], pnl2[0])); acc = cadd(acc, cmul(s2[57], pnl2[1]));
                    const cf ps = cconj(acc);
                    const float pa = cabsf_(ps);
                    for (int i = 0; i < 64; i++) {
                        const int d = L->binToDataNL[i];
                        if (d == 255) continue;
                        q0[d] = cdivs(cmul(s1[i], ps), pa); q1[d] = cdivs(cmul(s2[i], ps), pa);
                        inted[d] = fdiv(fadd(q0[d].re, q1[d].re), 2.0f);
                    }
                } else have = false;
                if (have) {
                    for (int i = 0; i < 52; i++) coded[L->deintNL[0][0][i]] = inted[i];
                    sig_viterbi(L, coded, sb, 26);
                    bcc_encode(sb, enc, 26);
                    double n0 = 0.0, n1 = 0.0;
                    for (int i = 0; i < 52; i++) {
                        const float ref = enc[L->deintNL[0][0][i]] ? 1.0f : -1.0f;
                        const cf e0 = mk(fsub(q0[i].re, ref), q0[i].im);
                        n0 += (double)fadd(fmul(e0.re, e0.re), fmul(e0.im, e0.im));
                        if (m.nSS == 2) { const cf e1 = mk(fsub(q1[i].re, ref), q1[i].im); n1 += (double)fadd(fmul(e1.re, e1.re), fmul(e1.im, e1.im)); }
                    }
                    sssnr0 = (float)(log10(52.0 / n0) * 10.0);
                    if (m.nSS == 2) sssnr1 = (float)(log10(52.0 / n1) * 10.0);
                } else {
                    for (int i = 0; i < 26; i++) sb[i] = 0;
                }
                parse_vhtb(sb, &m);
            }
            const int nl = (llen * 8 + 22 + 23) / 24;
            const bool ok = m.len > 0 && m.len <= 4095 && m.nSS <= 2 && (nl * 80) >= (m.nSym * m.nSymSamp + 160 + 80 + m.nLTF * 80 + 80);
            pos += need;
            if (!ok) return C8B_ST_FORMAT;
            trellis = m.nSym * m.nDBPS;
        } else {
            sig_viterbi(L, llrht, hb, 48);
            if (check_ht(hb)) {                                   // DEMOD_S_HT :180-216
                parse_ht(hb, &m);
                pos = 160;
                const int need = 80 + m.nLTF * 80;
                if (nsig - pos < need) return C8B_ST_TRUNC;
                chan_estimate(pos + 80);
                const int nl = (llen * 8 + 22 + 23) / 24;
                const bool ok = m.len > 0 && m.len <= 4095 && m.nSS <= 2 && (nl * 80) >= (m.nSym * m.nSymSamp + 160 + 80 + m.nLTF * 80);
                pos += need;
                if (!ok) return C8B_ST_FORMAT;
                trellis = m.len * 8 + 22;
            } else legacy = true;
        }
    }
    if (legacy) { parse_l(lmcs, llen, &m); trellis = m.len * 8 + 22; }
    f->format = m.format; f->mcs = m.mcs; f->len = m.len; f->cr = m.cr; f->ampdu = m.ampdu;
    f->nss = m.nSS; f->nsym = m.nSym; f->nsymsamp = m.nSymSamp; f->ncbps = m.nCBPS; f->ndbps = m.nDBPS;
    f->trellis = trellis; f->total = m.nSym * m.nCBPS; f->data_off = pos;
    f->sssnr0 = (m.format == C8B_F_VHT) ? sssnr0 : 0.f;
    f->sssnr1 = (m.format == C8B_F_VHT && m.nSS == 2) ? sssnr1 : 0.f;
    for (int i = 0; i < 64; i++) {
        const cf hh = (m.format == C8B_F_L) ? hl[i] : H[i][0];
        const bool used = (m.format == C8B_F_L) ? !l_null(i) : !nl_null(i);
        const double den = dadd(dmul((double)hh.re, (double)hh.re), dmul((double)hh.im, (double)hh.im));
        hinv[i] = (used && m.nSS == 1) ? mk((float)((double)hh.re / den), (float)(-(double)hh.im / den)) : mk(0.f, 0.f);
        for (int k = 0; k < 4; k++) w2[4 * i + k] = mk(0.f, 0.f);
        if (m.nSS == 2 && !nl_null(i)) {
            const cd h0 = dconj(tod(H[i][0])), h1 = dconj(tod(H[i][1])), h2 = dconj(tod(H[i][2])), h3 = dconj(tod(H[i][3]));
            const cd i0 = tod(HI[i][0]), i1 = tod(HI[i][1]), i2 = tod(HI[i][2]), i3 = tod(HI[i][3]);
            w2[4 * i + 0] = tof(dcadd(dcmul(h0, i0), dcmul(h2, i2)));   // stream 0 <- antenna 0
            w2[4 * i + 1] = tof(dcadd(dcmul(h1, i0), dcmul(h3, i2)));   // stream 0 <- antenna 1
            w2[4 * i + 2] = tof(dcadd(dcmul(h0, i1), dcmul(h2, i3)));   // stream 1 <- antenna 0
            w2[4 * i + 3] = tof(dcadd(dcmul(h1, i1), dcmul(h3, i3)));   // stream 1 <- antenna 1
        }
    }
    for (int q = 0; q < 4; q++) { w2[256 + q] = pnl[q]; w2[260 + q] = pnl2[q]; }
    if (m.nSS != 1 && m.nSS != 2) return C8B_ST_FORMAT;
    if (m.nSS == 1 && m.format != C8B_F_L && m.nLTF != 1) return C8B_ST_FORMAT;   // reference leaves H unset (:355-372)
    if (pos + m.nSym * m.nSymSamp > nsig) return C8B_ST_TRUNC;
    return C8B_ST_OK;
}

}  // namespace c8b
