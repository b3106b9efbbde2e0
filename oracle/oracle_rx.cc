// TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT (see oracle_rx.h for the rules and parity status).
//
// CPU restatement of the gr-ieee80211 20 MHz OFDM receive chain, frame/item level, float32 with
// the reference's operation order.  "ref:" comments give the reference file:line each function
// follows (paths relative to /root/reference).  Tables are generated from the 802.11 formulas
// (not transcribed) and compared with the reference's tables in tests/test_oracle_vs_ref.py.
#include "oracle_rx.h"

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstring>
#include <vector>
#include <atomic>
#include <thread>

typedef std::complex<float> cf;

namespace {

// ------------------------------------------------------------------------------------------------
// constants (ref: lib/cloud80211phy.h:29-56)
// ------------------------------------------------------------------------------------------------
enum { F_L = 0, F_HT = 1, F_VHT = 2 };
enum { CR_12 = 0, CR_23 = 1, CR_34 = 2, CR_56 = 3 };
enum { Q_BPSK = 0, Q_QBPSK = 1, Q_QPSK = 2, Q_16QAM = 3, Q_64QAM = 4, Q_256QAM = 5 };
const int SYM_SHIFT = 8;             // C8P_SYM_SAMP_SHIFT
const int DECODE_B_MAX = 4095;       // ref: lib/decode_impl.h:35
const int DECODE_T_MAX = 32782;      // ref: lib/decode_impl.h:36

struct Mod {                         // ref: lib/cloud80211phy.h:58-98 (SU subset)
    int format, sumu, ampdu, nSym, nSymSamp, nSD, nSP, nSS, nLTF;
    int mcs, len, mod, cr, nBPSCS, nDBPS, nCBPS, nCBPSS;
};

// ------------------------------------------------------------------------------------------------
// tables, by formula
// ------------------------------------------------------------------------------------------------
struct Tables {
    float ltfL[64], ltfNL[64], ltfNL22[64];   // FFT-bin order (bin k = +k, bin 64-k = -k)
    float pilotP[127];
    int deintL[4][288];                       // nBPSC 1,2,4,6
    int deintNL[2][5][416];                   // [iss-1][nBPSCS 1,2,4,6,8]
    int lsigDemap[64];                        // bin -> deinterleaved L-SIG llr index, -1 unused
    int nxt[64][2], outp[64][2];
    uint32_t crc32tab[256];
    Tables()
    {
        // IEEE 802.11-2016 Eq. (17-8): L-LTF on subcarriers -26..26
        static const int8_t L[53] = { 1, 1, -1, -1, 1, 1, -1, 1, -1, 1, 1, 1, 1, 1, 1, -1, -1, 1, 1, -1, 1, -1, 1, 1, 1, 1,
                                      0, 1, -1, -1, 1, 1, -1, 1, -1, 1, -1, -1, -1, -1, -1, 1, 1, -1, -1, 1, -1, 1, -1, 1, 1, 1, 1 };
        for (int i = 0; i < 64; i++) { ltfL[i] = 0.f; ltfNL[i] = 0.f; }
        for (int k = -26; k <= 26; k++) { ltfL[(k + 64) & 63] = (float)L[k + 26]; ltfNL[(k + 64) & 63] = (float)L[k + 26]; }
        // Eq. (19-23): HT-LTF adds {1,1} at -28,-27 and {-1,-1} at 27,28
        ltfNL[(-28 + 64) & 63] = 1.f; ltfNL[(-27 + 64) & 63] = 1.f; ltfNL[27] = -1.f; ltfNL[28] = -1.f;
        // second VHT-LTF of a 2-LTF frame seen through P row 2: data tones negated, pilot tones not
        // (R-matrix), i.e. relative to -LTF the four pilot bins flip.  ref table: LTF_NL_28_F_FLOAT_VHT22.
        for (int i = 0; i < 64; i++) ltfNL22[i] = ltfNL[i];
        ltfNL22[7] = -ltfNL[7]; ltfNL22[21] = -ltfNL[21]; ltfNL22[43] = -ltfNL[43]; ltfNL22[57] = -ltfNL[57];

        // pilot polarity p_0..126: scrambler x^7+x^4+1 seeded all ones, 0 -> +1, 1 -> -1 (17.3.5.10)
        int st = 0x7f;
        for (int i = 0; i < 127; i++) {
            int fb = ((st >> 6) ^ (st >> 3)) & 1;
            st = ((st << 1) & 0x7e) | fb;
            pilotP[i] = fb ? -1.f : 1.f;
        }
        // legacy interleaver (17.3.5.7): k -> i -> j ; deinterleave scatter map[j] = k
        const int nb[4] = { 1, 2, 4, 6 };
        for (int m = 0; m < 4; m++) {
            int ncbps = 48 * nb[m], s = std::max(nb[m] / 2, 1);
            for (int k = 0; k < ncbps; k++) {
                int i = (ncbps / 16) * (k % 16) + k / 16;
                int j = s * (i / s) + (i + ncbps - (16 * i) / ncbps) % s;
                deintL[m][j] = k;
            }
        }
        // HT/VHT 20 MHz interleaver (19.3.11.8.3): N_COL 13, N_ROW 4*N_BPSCS, N_ROT 11
        const int nbn[5] = { 1, 2, 4, 6, 8 };
        for (int iss = 1; iss <= 2; iss++)
            for (int m = 0; m < 5; m++) {
                int ncbpss = 52 * nbn[m], s = std::max(nbn[m] / 2, 1), nrow = 4 * nbn[m];
                for (int k = 0; k < ncbpss; k++) {
                    int i = nrow * (k % 13) + k / 13;
                    int j = s * (i / s) + (i + ncbpss - (13 * i) / ncbpss) % s;
                    int rot = (((iss - 1) * 2) % 3 + 3 * ((iss - 1) / 3)) * 11 * nbn[m];
                    int r = ((j - rot) % ncbpss + ncbpss) % ncbpss;
                    deintNL[iss - 1][m][r] = k;
                }
            }
        // L-SIG / HT-SIG / VHT-SIG-A tone -> llr index: data tone d (in -26..26 order) -> deintL[BPSK][d]
        int d = 0;
        for (int i = 0; i < 64; i++) lsigDemap[i] = -1;
        for (int k = -26; k <= 26; k++) {
            if (k == 0 || k == -21 || k == -7 || k == 7 || k == 21) continue;
            lsigDemap[(k + 64) & 63] = deintL[0][d++];
        }
        // trellis (ref: c8p.cc:1864-1887, generated from g0=133o, g1=171o as in bccEncoder c8p.cc:2622-2644)
        for (int s6 = 0; s6 < 64; s6++)
            for (int b = 0; b < 2; b++) {
                nxt[s6][b] = (s6 >> 1) | (b << 5);
                int reg = b;                                  // 7-bit register, newest bit at LSB
                for (int q = 0; q < 6; q++) reg |= ((s6 >> (5 - q)) & 1) << (q + 1);
                int o0 = __builtin_popcount(reg & 0155) & 1, o1 = __builtin_popcount(reg & 0117) & 1;
                outp[s6][b] = o0 * 2 + o1;
            }
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c & 1) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
            crc32tab[i] = c;
        }
    }
};
const Tables T;

int nbIndexNL(int nbpscs) { return nbpscs == 1 ? 0 : nbpscs == 2 ? 1 : nbpscs == 4 ? 2 : nbpscs == 6 ? 3 : nbpscs == 8 ? 4 : -1; }
int nbIndexL(int nbpsc) { return nbpsc == 1 ? 0 : nbpsc == 2 ? 1 : nbpsc == 4 ? 2 : nbpsc == 6 ? 3 : -1; }

// ------------------------------------------------------------------------------------------------
// FFT (stand-in for gr::fft::fft_complex_fwd(64): forward, unnormalised, natural order;
// call sites lib/signal_impl.cc:121-123, lib/demod_impl.cc:541-547).  UNPINNED: FFTW is absent;
// evaluated in double with a radix-2 DIT and rounded once, i.e. the correctly rounded DFT up to
// double round-off, which any float FFT (FFTW's included) approximates to ~1e-7 relative.
// ------------------------------------------------------------------------------------------------
struct Tw { double c[32], s[32]; int rev[64]; Tw() { for (int i = 0; i < 32; i++) { c[i] = cos(-2.0 * M_PI * i / 64.0); s[i] = sin(-2.0 * M_PI * i / 64.0); }
        for (int i = 0; i < 64; i++) { int r = 0; for (int b = 0; b < 6; b++) r |= ((i >> b) & 1) << (5 - b); rev[i] = r; } } };
const Tw TW;

void fft64(const cf* in, cf* out)
{
    double re[64], im[64];
    for (int i = 0; i < 64; i++) { re[TW.rev[i]] = in[i].real(); im[TW.rev[i]] = in[i].imag(); }
    for (int len = 2; len <= 64; len <<= 1) {
        int half = len >> 1, step = 64 / len;
        for (int b = 0; b < 64; b += len)
            for (int k = 0; k < half; k++) {
                double wr = TW.c[k * step], wi = TW.s[k * step];
                double xr = re[b + k + half] * wr - im[b + k + half] * wi;
                double xi = re[b + k + half] * wi + im[b + k + half] * wr;
                re[b + k + half] = re[b + k] - xr; im[b + k + half] = im[b + k] - xi;
                re[b + k] += xr; im[b + k] += xi;
            }
    }
    for (int i = 0; i < 64; i++) out[i] = cf((float)re[i], (float)im[i]);
}

// ------------------------------------------------------------------------------------------------
// presiso (ref: examples/presiso.grc:35-229; stock GNU Radio blocks delay(16),
// multiply_conjugate_cc, moving_average_cc(48), complex_to_mag, complex_to_mag_squared,
// moving_average_ff(64), divide_ff).  UNPINNED summation order: GNU Radio's moving_average keeps a
// running sum that is re-seeded every work() call / 4000 items, so its rounding depends on the
// scheduler.  The oracle fixes this binary tree over the window ending at n (v[m]=0 for m<0):
//   s2[n]=v[n-1]+v[n]  s4[n]=s2[n-2]+s2[n]  s8[n]=s4[n-4]+s4[n]  s16[n]=s8[n-8]+s8[n]
//   sum48[n]=(s16[n-32]+s16[n-16])+s16[n]     sum64[n]=(s16[n-48]+s16[n-32])+(s16[n-16]+s16[n])
// ------------------------------------------------------------------------------------------------
template <class V> void tree16(const std::vector<V>& v, std::vector<V>& s16)
{
    int n = (int)v.size();
    std::vector<V> a(n), b(n);
    auto at = [](const std::vector<V>& x, int i) { return i >= 0 ? x[i] : V(0); };
    for (int i = 0; i < n; i++) a[i] = at(v, i - 1) + v[i];
    for (int i = 0; i < n; i++) b[i] = at(a, i - 2) + a[i];
    for (int i = 0; i < n; i++) a[i] = at(b, i - 4) + b[i];
    s16.resize(n);
    for (int i = 0; i < n; i++) s16[i] = at(a, i - 8) + a[i];
}

void presiso(const cf* x, int n, float* preac, cf* preconj)
{
    std::vector<cf> p(n), p16;
    std::vector<float> a(n), a16;
    for (int i = 0; i < n; i++) {
        cf d = i >= 16 ? x[i - 16] : cf(0.f, 0.f);               // blocks.delay(16)
        // multiply_conjugate_cc: in0 * conj(in1), in0 = delayed (presiso.grc:220,228)
        float re = d.real() * x[i].real() + d.imag() * x[i].imag();
        float im = d.imag() * x[i].real() - d.real() * x[i].imag();
        p[i] = cf(re, im);
        a[i] = x[i].real() * x[i].real() + x[i].imag() * x[i].imag();   // complex_to_mag_squared
    }
    tree16(p, p16);
    tree16(a, a16);
    auto pc = [&](int i) { return i >= 0 ? p16[i] : cf(0.f, 0.f); };
    auto pa = [&](int i) { return i >= 0 ? a16[i] : 0.f; };
    for (int i = 0; i < n; i++) {
        cf c = (pc(i - 32) + pc(i - 16)) + pc(i);                      // moving_average_cc(48)
        float pw = (pa(i - 48) + pa(i - 32)) + (pa(i - 16) + pa(i));   // moving_average_ff(64)
        float mag = sqrtf(c.real() * c.real() + c.imag() * c.imag()); // complex_to_mag
        preconj[i] = c;
        preac[i] = mag / pw;                                           // divide_ff (0/0 -> NaN, as GR)
    }
}

// ------------------------------------------------------------------------------------------------
// trigger (ref: lib/trigger_impl.cc:59-117)
// ------------------------------------------------------------------------------------------------
struct TrigState { int nPlateau, fPlateau, fPlateauEnd, countDown; float conjAc; };

void trigger(const float* ac, int n, uint8_t* out, TrigState& s)
{
    for (int i = 0; i < n; i++) {
        uint8_t o = 0;
        if (ac[i] > 0.3f) {                                   // :79
            s.nPlateau++;
            if (ac[i] > s.conjAc) { s.conjAc = ac[i]; o |= 0x02; }      // :82-87
            if (s.nPlateau > 20 && (s.fPlateau + s.fPlateauEnd) == 0) { // :88-93
                s.fPlateau = 1; s.fPlateauEnd = 1; s.countDown = 80;
            }
        } else {                                              // :95-100
            s.nPlateau = 0; s.fPlateauEnd = 0; s.conjAc = 0.0f;
        }
        if (s.fPlateau) {                                     // :101-109
            s.countDown--;
            if (s.countDown == 0) { s.fPlateau = 0; o |= 0x01; }
        }
        out[i] = o;
    }
}

// ------------------------------------------------------------------------------------------------
// sync (ref: lib/sync_impl.cc:92-147 SYNC state, :155-179 ltf_autoCorrelation, :181-196 ltf_cfo)
// ------------------------------------------------------------------------------------------------
const int SYNC_BUF = 240, SYNC_RES = 111;    // ref: lib/sync_impl.h:30-31

struct SyncOut { int ok, mIndex; float rad, snr, rssi; };

SyncOut syncAt(const cf* sig, cf conjAvg, float* acOut)
{
    float ac[SYNC_RES], pwr[SYNC_RES];
    cf msum(0.f, 0.f);
    float s1 = 0.f, s2 = 0.f;
    for (int i = 0; i < 64; i++) {                            // :161-166
        msum += sig[i] * std::conj(sig[i + 64]);
        s1 += std::abs(sig[i]) * std::abs(sig[i]);
        s2 += std::abs(sig[i + 64]) * std::abs(sig[i + 64]);
    }
    for (int i = 0; i < SYNC_RES; i++) {                      // :167-178
        ac[i] = std::abs(msum) / std::sqrt(s1) / std::sqrt(s2);
        pwr[i] = s1;
        msum -= sig[i] * std::conj(sig[i + 64]);
        s1 -= std::abs(sig[i]) * std::abs(sig[i]);
        s2 -= std::abs(sig[i + 64]) * std::abs(sig[i + 64]);
        msum += sig[i + 64] * std::conj(sig[i + 128]);
        s1 += std::abs(sig[i + 64]) * std::abs(sig[i + 64]);
        s2 += std::abs(sig[i + 128]) * std::abs(sig[i + 128]);
    }
    if (acOut) memcpy(acOut, ac, sizeof(ac));
    SyncOut o; o.ok = 0; o.mIndex = 0; o.rad = o.snr = o.rssi = 0.f;
    float* mp = std::max_element(ac, ac + SYNC_RES);           // :97 (first maximum)
    if (*mp > 0.5) {                                          // :99
        float thr = *mp * 0.8;                                // :101 float*double -> float
        double maxD = (double)(*mp);
        int mi = (int)(mp - ac), l = mi, r = mi;
        for (int j = mi; j >= 0; j--) if (ac[j] < thr) { l = j; break; }           // :105-112
        for (int j = mi; j < SYNC_RES; j++) if (ac[j] < thr) { r = j; break; }     // :113-120
        o.ok = 1; o.mIndex = (l + r) / 2;
        // ltf_cfo(&inSig[mIndex]) :181-196
        const cf* s = sig + o.mIndex;
        float radStf = atan2f(conjAvg.imag(), conjAvg.real()) / 16.0f;
        cf tmp[128];
        for (int i = 0; i < 128; i++) tmp[i] = s[i] * cf(cosf(i * radStf), sinf(i * radStf));
        cf csum(0.f, 0.f);
        for (int i = 0; i < 64; i++) csum += tmp[i] * std::conj(tmp[i + 64]);
        float radLtf = atan2f((csum / 64.0f).imag(), (csum / 64.0f).real()) / 64.0f;
        o.rad = radStf + radLtf;
        o.snr = (float)(10.0 * log10(maxD / (1.0 - maxD)));   // :126
        o.rssi = pwr[mi] / 64.0f;                             // :127
    }
    return o;
}

// ------------------------------------------------------------------------------------------------
// SIG-field Viterbi, trellis <= 48 (ref: c8p.cc:2001-2088 svSigDecoder::decode) and the decode
// block's Viterbi (ref: lib/decode_impl.cc:164-302) share this ACS; `his` keeps the reference's
// "previous state, default 0" encoding (one byte per state per step).
// ------------------------------------------------------------------------------------------------
inline void acsStep(const float* pre, float* cur, uint8_t* hisRow, float t0, float t1)
{
    float tab[4] = { 0.0f, t1, t0, t1 + t0 };                 // decode_impl.cc:231-234
    for (int i = 0; i < 64; i++) { cur[i] = -1000000000000000.0f; hisRow[i] = 0; }
    for (int i = 0; i < 64; i++) {                            // decode_impl.cc:237-260
        float a0 = pre[i] + tab[T.outp[i][0]];
        float a1 = pre[i] + tab[T.outp[i][1]];
        int n0 = T.nxt[i][0], n1 = T.nxt[i][1];
        if (a0 > cur[n0]) { cur[n0] = a0; hisRow[n0] = (uint8_t)i; }
        if (a1 > cur[n1]) { cur[n1] = a1; hisRow[n1] = (uint8_t)i; }
    }
}

void traceback(const uint8_t* his /* [(trellis+1)][64] */, int trellis, uint8_t* bits)
{
    std::vector<uint8_t> seq(trellis + 1);
    seq[trellis] = 0;                                         // decode_impl.cc:285 final state 0
    for (int j = trellis; j > 0; j--) seq[j - 1] = his[(size_t)j * 64 + seq[j]];
    for (int j = 0; j < trellis; j++) bits[j] = (seq[j + 1] == T.nxt[seq[j]][1]) ? 1 : 0;
}

void sigViterbi(const float* llr, uint8_t* bits, int trellis)
{
    if (trellis < 0 || trellis > 48) return;                  // c8p.cc:2003
    float m0[64], m1[64];
    uint8_t his[49 * 64];
    memset(his, 0, sizeof(his));
    for (int i = 0; i < 64; i++) m0[i] = -1000000000000000.0f;
    m0[0] = 0;
    float *pre = m0, *cur = m1;
    for (int t = 0; t < trellis; t++) {
        acsStep(pre, cur, his + (size_t)(t + 1) * 64, llr[2 * t], llr[2 * t + 1]);
        std::swap(pre, cur);
    }
    traceback(his, trellis, bits);
}

// CRC-8 of HT-SIG / VHT-SIG-A (ref: c8p.cc:1367-1403 checkBitCrc8)
bool crc8Check(const uint8_t* bits, int len, const uint8_t* crc)
{
    unsigned c = 0xff;
    for (int i = 0; i < len; i++) {
        unsigned top = (c >> 7) & 1;
        c = (c << 1) & 0xff;
        if (top) c ^= 0x07;
        if (bits[i]) c ^= 0x07;
    }
    for (int i = 0; i < 8; i++) if (crc[i]) c ^= (1u << (7 - i));
    return (c & 0xff) == 0xff;
}

// ------------------------------------------------------------------------------------------------
// L-SIG (ref: c8p.cc:609-627 procLHSigDemodDeint, :650-728 signalCheckLegacy)
// ------------------------------------------------------------------------------------------------
void lsigDemod(const cf* s1, const cf* s2, const cf* sig, cf* h, float* llr)
{
    const int pb[4] = { 7, 21, 43, 57 };
    for (int q = 0; q < 4; q++) h[pb[q]] = (s1[pb[q]] + s2[pb[q]]) / (2.0f * T.ltfL[pb[q]]);
    cf ps = std::conj(sig[7] / h[7] - sig[21] / h[21] + sig[43] / h[43] + sig[57] / h[57]);
    float pa = std::abs(ps);
    for (int i = 0; i < 64; i++)
        if (T.lsigDemap[i] > -1) {
            h[i] = (s1[i] + s2[i]) / (2.0f * T.ltfL[i]);
            llr[T.lsigDemap[i]] = (sig[i] / h[i] * ps / pa).real();
        }
}

const int L_NDBPS[8] = { 24, 36, 48, 72, 96, 144, 192, 216 };

bool lsigCheck(const uint8_t* b, int* mcs, int* len, int* ndbps)
{
    if (!b[3]) return false;
    if (b[4]) return false;
    int par = 0;
    for (int i = 0; i < 17; i++) par += b[i];
    if ((par & 1) ^ b[17]) return false;
    int rate = b[0] | (b[1] << 1) | (b[2] << 2) | (b[3] << 3);
    // R1-R4 (bit0 first) -> rate index; rate=8+... since b[3]==1
    static const int rmap[16] = { 0, 0, 0, 0, 0, 0, 0, 0, 6, 4, 2, 0, 7, 5, 3, 1 };
    *mcs = rmap[rate];
    *ndbps = L_NDBPS[*mcs];
    *len = 0;
    for (int i = 0; i < 12; i++) *len |= ((int)b[i + 5]) << i;
    if (*len > 4095 || *len < 14) return false;
    return true;
}

// signal block S_DEMOD (ref: lib/signal_impl.cc:108-162)
int signalAt(const cf* in, float rad, cf* h, float* llr, uint8_t* bits, int* mcs, int* len, int* nsamp)
{
    cf f1[64], f2[64], fs[64], o1[64], o2[64], os[64];
    for (int i = 0; i < 64; i++) {                            // :115-120
        f1[i] = in[SYM_SHIFT + i] * cf(cosf((i + SYM_SHIFT) * rad), sinf((i + SYM_SHIFT) * rad));
        f2[i] = in[SYM_SHIFT + 64 + i] * cf(cosf((i + SYM_SHIFT + 64) * rad), sinf((i + SYM_SHIFT + 64) * rad));
        fs[i] = in[SYM_SHIFT + 144 + i] * cf(cosf((i + SYM_SHIFT + 144) * rad), sinf((i + SYM_SHIFT + 144) * rad));
    }
    fft64(f1, o1); fft64(f2, o2); fft64(fs, os);
    for (int i = 0; i < 64; i++) h[i] = cf(0.f, 0.f);
    lsigDemod(o1, o2, os, h, llr);
    sigViterbi(llr, bits, 24);
    int ndbps = 24;
    if (!lsigCheck(bits, mcs, len, &ndbps)) return 0;
    int nsym = (*len * 8 + 22 + ndbps - 1) / ndbps;           // :128
    *nsamp = nsym * 80;
    return 1;
}

// ------------------------------------------------------------------------------------------------
// SIG parsers (ref: c8p.cc:773-850 signalParserL, :852-998 signalParserHt, :1090-1178
// signalParserVhtA, :1180-1223 signalParserVhtB, :1225-1323 modParserVht, :730-771 checks)
// ------------------------------------------------------------------------------------------------
int bitsToInt(const uint8_t* b, int n) { int v = 0; for (int i = 0; i < n; i++) v |= ((int)b[i]) << i; return v; }

void rateFields(Mod* m)     // shared tail of modParserHt / modParserVht / signalParserHt
{
    m->nSD = 52; m->nSP = 4;
    m->nCBPSS = m->nBPSCS * m->nSD;
    m->nCBPS = m->nCBPSS * m->nSS;
    switch (m->cr) {
    case CR_12: m->nDBPS = m->nCBPS / 2; break;
    case CR_23: m->nDBPS = (m->nCBPS * 2) / 3; break;
    case CR_34: m->nDBPS = (m->nCBPS * 3) / 4; break;
    case CR_56: m->nDBPS = (m->nCBPS * 5) / 6; break;
    }
    switch (m->nSS) { case 1: m->nLTF = 1; break; case 2: m->nLTF = 2; break; case 3: case 4: m->nLTF = 4; break; default: break; }
}

void parseL(int mcs, int len, Mod* m)
{
    static const int md[8] = { Q_BPSK, Q_BPSK, Q_QPSK, Q_QPSK, Q_16QAM, Q_16QAM, Q_64QAM, Q_64QAM };
    static const int cr[8] = { CR_12, CR_34, CR_12, CR_34, CR_12, CR_34, CR_23, CR_34 };
    static const int nb[8] = { 1, 1, 2, 2, 4, 4, 6, 6 };
    m->mcs = mcs;
    if (mcs >= 0 && mcs < 8) { m->mod = md[mcs]; m->cr = cr[mcs]; m->nBPSCS = nb[mcs]; m->nDBPS = L_NDBPS[mcs]; m->nCBPS = 48 * nb[mcs]; }
    m->len = len; m->nCBPSS = m->nCBPS; m->nSD = 48; m->nSP = 4; m->nSS = 1; m->sumu = 0; m->nLTF = 0;
    m->format = F_L; m->nSymSamp = 80;
    m->nSym = (len * 8 + 22) / m->nDBPS + (((len * 8 + 22) % m->nDBPS) != 0);
    m->ampdu = 0;
}

void modHtIndex(int mcs8, Mod* m)   // HT mcs%8 -> modulation/rate (c8p.cc:909-953)
{
    static const int md[8] = { Q_BPSK, Q_QPSK, Q_QPSK, Q_16QAM, Q_16QAM, Q_64QAM, Q_64QAM, Q_64QAM };
    static const int nb[8] = { 1, 2, 2, 4, 4, 6, 6, 6 };
    static const int cr[8] = { CR_12, CR_12, CR_34, CR_12, CR_34, CR_23, CR_34, CR_56 };
    m->mod = md[mcs8]; m->nBPSCS = nb[mcs8]; m->cr = cr[mcs8];
}

bool checkHt(const uint8_t* b)
{
    if (b[26] != 1) return false;
    if (!crc8Check(b, 34, b + 34)) return false;
    if (b[5] + b[6] + b[7] + b[28] + b[29] + b[30] + b[32] + b[33]) return false;
    return true;
}

bool checkVhtA(const uint8_t* b)
{
    if (b[2] != 1 || b[23] != 1 || b[33] != 1) return false;
    if (!crc8Check(b, 34, b + 34)) return false;
    if (b[0] + b[1]) return false;
    return true;
}

void parseHt(const uint8_t* b, Mod* m)
{
    int mcs = bitsToInt(b, 7), len = bitsToInt(b + 8, 16);
    m->format = F_HT; m->sumu = 0;
    m->nSymSamp = b[31] ? 72 : 80;
    m->ampdu = b[27] ? 1 : 0;
    m->mcs = mcs;
    modHtIndex(mcs % 8, m);
    m->len = len;
    m->nSS = mcs / 8 + 1;
    rateFields(m);
    m->nSym = (len * 8 + 22) / m->nDBPS + (((len * 8 + 22) % m->nDBPS) != 0);
}

void modVht(int mcs, Mod* m)
{
    static const int md[10] = { Q_BPSK, Q_QPSK, Q_QPSK, Q_16QAM, Q_16QAM, Q_64QAM, Q_64QAM, Q_64QAM, Q_256QAM, Q_256QAM };
    static const int nb[10] = { 1, 2, 2, 4, 4, 6, 6, 6, 8, 8 };
    static const int cr[10] = { CR_12, CR_12, CR_34, CR_12, CR_34, CR_23, CR_34, CR_56, CR_34, CR_56 };
    m->mcs = mcs;
    if (mcs >= 0 && mcs < 10) { m->mod = md[mcs]; m->nBPSCS = nb[mcs]; m->cr = cr[mcs]; }   // else: fields keep old values (c8p.cc:1280)
    rateFields(m);
}

void parseVhtA(const uint8_t* b, Mod* m)
{
    int gid = bitsToInt(b + 4, 6);
    m->format = F_VHT;
    m->nSymSamp = b[24] ? 72 : 80;
    m->ampdu = 1;
    if (gid == 0 || gid == 63) {
        m->sumu = 0;
        m->nSS = bitsToInt(b + 10, 3) + 1;
        modVht(bitsToInt(b + 28, 4), m);
    } else {
        m->sumu = 1; m->nLTF = 2; m->nSS = 1; m->nSD = 52; m->nSP = 4;
    }
}

void parseVhtB(const uint8_t* b, Mod* m)
{
    // An MCS field outside 0..9 leaves nDBPS at whatever the block's member held (c8p.cc:1225-1294 default case): the
    // reference then divides by a stale value, or by zero on a fresh block.  Documented delta: such a frame is dropped
    // at the sanity check (len = nSym = -1).
    if (m->sumu) {
        int len = bitsToInt(b, 16), mcs = bitsToInt(b + 16, 4);
        modVht(mcs, m);
        m->nLTF = 2;
        if (m->nDBPS <= 0) { m->len = -1; m->nSym = -1; return; }
        m->len = len * 4;
        m->nSym = (m->len * 8 + 16 + 6) / m->nDBPS + (((m->len * 8 + 16 + 6) % m->nDBPS) != 0);
        m->nLTF = 2;
    } else if ((b[17] + b[18] + b[19]) == 3) {
        if (m->nDBPS <= 0) { m->len = -1; m->nSym = -1; return; }
        m->len = bitsToInt(b, 17) * 4;
        m->nSym = (m->len * 8 + 16 + 6) / m->nDBPS + (((m->len * 8 + 16 + 6) % m->nDBPS) != 0);
    } else {
        unsigned pat = (unsigned)bitsToInt(b, 20);
        if (pat == 0b01000010001011100000u) { m->len = 0; m->nSym = 0; }   // NDP
        else { m->len = -1; m->nSym = -1; }
    }
}

// BCC encoder (ref: c8p.cc:2622-2644), used by VHT-SIG-B SNR estimate
void bcc(const uint8_t* in, uint8_t* out, int len)
{
    int st = 0;
    for (int i = 0; i < len; i++) {
        st = ((st << 1) & 0x7e) | in[i];
        out[2 * i] = __builtin_popcount(st & 0155) & 1;
        out[2 * i + 1] = __builtin_popcount(st & 0117) & 1;
    }
}

// ------------------------------------------------------------------------------------------------
// LLR demap (ref: c8p.cc:2090-2148 procSymQamToLlr); qam is scaled in place like the reference
// ------------------------------------------------------------------------------------------------
void qamToLlr(cf* q, float* o, int mod, int nsd)
{
    if (mod == Q_BPSK) { for (int i = 0; i < nsd; i++) o[i] = q[i].real(); }
    else if (mod == Q_QPSK) { for (int i = 0; i < nsd; i++) { q[i] *= 1.4142135623730951f; o[2 * i] = q[i].real(); o[2 * i + 1] = q[i].imag(); } }
    else if (mod == Q_16QAM) {
        for (int i = 0; i < nsd; i++) {
            q[i] *= 3.1622776601683795f;
            o[4 * i] = q[i].real(); o[4 * i + 1] = 2.0f - std::abs(q[i].real());
            o[4 * i + 2] = q[i].imag(); o[4 * i + 3] = 2.0f - std::abs(q[i].imag());
        }
    } else if (mod == Q_64QAM) {
        for (int i = 0; i < nsd; i++) {
            q[i] *= 6.48074069840786f;
            float r = q[i].real(), m = q[i].imag();
            o[6 * i] = r; o[6 * i + 1] = 4.0f - std::abs(r); o[6 * i + 2] = 2 - std::abs(4.0f - std::abs(r));
            o[6 * i + 3] = m; o[6 * i + 4] = 4.0f - std::abs(m); o[6 * i + 5] = 2 - std::abs(4.0f - std::abs(m));
        }
    } else if (mod == Q_256QAM) {
        for (int i = 0; i < nsd; i++) {
            q[i] *= 13.038404810405298f;
            float r = q[i].real(), m = q[i].imag();
            o[8 * i] = r; o[8 * i + 1] = 8.0f - std::abs(r); o[8 * i + 2] = 4 - std::abs(8.0f - std::abs(r));
            o[8 * i + 3] = 2 - std::abs(4 - std::abs(8.0f - std::abs(r)));
            o[8 * i + 4] = m; o[8 * i + 5] = 8.0f - std::abs(m); o[8 * i + 6] = 4 - std::abs(8.0f - std::abs(m));
            o[8 * i + 7] = 2 - std::abs(4 - std::abs(8.0f - std::abs(m)));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// demod (ref: lib/demod_impl.cc:59-342 state machine, :344-557 helpers)
// ------------------------------------------------------------------------------------------------
int g_mupos = 0;

struct Demod {
    Mod m;
    cf HL[64], HNL[64], sig1[64], qam[52], fo1[64], fo2[64];
    float pilot[4];
    int pilotP;
    float sssnr;
    cf mu2x1[128];

    void fftDemod(const cf* s, cf* o) { fft64(s, o); }        // :541-547
    static bool nlNull(int i) { return i == 0 || (i >= 29 && i <= 35); }
    static bool lNull(int i) { return i == 0 || (i >= 27 && i <= 37); }
    static bool isPilot(int i) { return i == 7 || i == 21 || i == 43 || i == 57; }

    void nonLegacyChanEstimate(const cf* s)                   // :344-411
    {
        if (m.format == F_VHT && m.sumu) {
            fftDemod(s + SYM_SHIFT, fo1); fftDemod(s + SYM_SHIFT + 80, fo2);
            for (int i = 0; i < 64; i++) {
                if (nlNull(i)) continue;
                if (g_mupos == 0) HNL[i] = (fo1[i] - fo2[i]) / (T.ltfNL[i] * 2.0f);
                else HNL[i] = (fo1[i] / T.ltfNL[i] + fo2[i] / T.ltfNL22[i]) / 2.0f;
            }
        } else if (m.nSS == 1 && m.nLTF == 1) {
            fftDemod(s + SYM_SHIFT, fo1);
            for (int i = 0; i < 64; i++) if (!nlNull(i)) HNL[i] = fo1[i] / T.ltfNL[i];
        } else {
            memcpy(&mu2x1[0], s + SYM_SHIFT, sizeof(cf) * 64);
            memcpy(&mu2x1[64], s + SYM_SHIFT + 80, sizeof(cf) * 64);
            fftDemod(s + SYM_SHIFT, fo1);
            for (int i = 0; i < 64; i++) if (!nlNull(i)) HNL[i] = fo1[i] / T.ltfNL[i];
        }
    }

    void pilotShift() { float t = pilot[0]; pilot[0] = pilot[1]; pilot[1] = pilot[2]; pilot[2] = pilot[3]; pilot[3] = t; }   // :549-557

    void nonLegacyChanUpdate(const cf* s)                     // :413-447
    {
        fftDemod(s + SYM_SHIFT, fo1);
        for (int i = 0; i < 64; i++) if (!nlNull(i)) sig1[i] = fo1[i] / HNL[i];
        cf ps = std::conj(sig1[7] * pilot[2] * T.pilotP[pilotP] + sig1[21] * pilot[3] * T.pilotP[pilotP] +
                          sig1[43] * pilot[0] * T.pilotP[pilotP] + sig1[57] * pilot[1] * T.pilotP[pilotP]);
        pilotShift();
        pilotP = (pilotP + 1) % 127;
        float pa = std::abs(ps);
        int j = 26;
        for (int i = 0; i < 64; i++) {
            if (nlNull(i) || isPilot(i)) continue;
            qam[j] = sig1[i] * ps / pa;
            if (++j >= 52) j = 0;
        }
    }

    void legacyChanUpdate(const cf* s)                        // :507-539
    {
        fftDemod(s + SYM_SHIFT, fo1);
        for (int i = 0; i < 64; i++) if (!lNull(i)) sig1[i] = fo1[i] / HL[i];
        cf ps = std::conj(sig1[7] * pilot[2] * T.pilotP[pilotP] + sig1[21] * pilot[3] * T.pilotP[pilotP] +
                          sig1[43] * pilot[0] * T.pilotP[pilotP] + sig1[57] * pilot[1] * T.pilotP[pilotP]);
        pilotP = (pilotP + 1) % 127;
        float pa = std::abs(ps);
        int j = 24;
        for (int i = 0; i < 64; i++) {
            if (lNull(i) || isPilot(i)) continue;
            qam[j] = sig1[i] * ps / pa;
            if (++j >= 48) j = 0;
        }
    }

    void vhtSigBDemod(const cf* s, uint8_t* bits26)           // :449-505
    {
        cf bq[52];
        float inted[52], coded[52];
        fftDemod(s + SYM_SHIFT, fo1);
        for (int i = 0; i < 64; i++) if (!nlNull(i)) sig1[i] = fo1[i] / HNL[i];
        cf ps = std::conj(sig1[7] - sig1[21] + sig1[43] + sig1[57]);
        float pa = std::abs(ps);
        int j = 26;
        for (int i = 0; i < 64; i++) {
            if (nlNull(i) || isPilot(i)) continue;
            bq[j] = sig1[i] * ps / pa;
            inted[j] = bq[j].real();
            if (++j >= 52) j = 0;
        }
        for (int i = 0; i < 52; i++) coded[T.deintNL[0][0][i]] = inted[i];     // mapDeintVhtSigB20
        sigViterbi(coded, bits26, 26);
        uint8_t enc[52], intl[52];
        bcc(bits26, enc, 26);
        for (int i = 0; i < 52; i++) intl[i] = enc[T.deintNL[0][0][i]];        // procIntelVhtB20: out[mapIntel[k]] = in[k]
        double np = 0.0;
        for (int i = 0; i < 52; i++) {
            if (intl[i]) bq[i] -= cf(1.0f, 0.0f); else bq[i] -= cf(-1.0f, 0.0f);
            np += (double)(bq[i].real() * bq[i].real() + bq[i].imag() * bq[i].imag());
        }
        sssnr = (float)(log10(52.0 / np) * 10.0);
    }
};

// HT-SIG / VHT-SIG-A demod (ref: c8p.cc:629-648 procNLSigDemodDeint)
void nlsigDemod(const cf* s1, const cf* s2, const cf* h, float* llrht, float* llrvht)
{
    cf p1 = std::conj(s1[7] / h[7] - s1[21] / h[21] + s1[43] / h[43] + s1[57] / h[57]);
    cf p2 = std::conj(s2[7] / h[7] - s2[21] / h[21] + s2[43] / h[43] + s2[57] / h[57]);
    float a1 = std::abs(p1), a2 = std::abs(p2);
    for (int i = 0; i < 64; i++)
        if (T.lsigDemap[i] > -1) {
            cf m1 = s1[i] / h[i] * p1 / a1;
            cf m2 = s2[i] / h[i] * p2 / a2;
            llrht[T.lsigDemap[i]] = m1.imag();
            llrht[T.lsigDemap[i] + 48] = m2.imag();
            llrvht[T.lsigDemap[i]] = m1.real();
            llrvht[T.lsigDemap[i] + 48] = m2.imag();
        }
}

// One frame through the demod state machine.  sig[0] is the first sample signal copies out
// (frame sample 400); nsig is how many valid samples follow (nsamp, +pad).  Returns status.
int demodFrame(const cf* sig, int nsig, int lmcs, int llen, const cf* hl, orx_frame* f, float* llrOut, int llrCap)
{
    Demod d;
    memset(&d.m, 0, sizeof(d.m));
    for (int i = 0; i < 64; i++) { d.HL[i] = hl[i]; d.HNL[i] = cf(0.f, 0.f); d.sig1[i] = cf(0.f, 0.f); }
    d.sssnr = 0.f;
    int pos = 0;                                              // samples consumed so far
    bool legacy = lmcs > 0;                                   // :93-100
    int trellis = 0;
    if (!legacy) {                                            // DEMOD_S_FORMAT :106-148
        if (nsig < 160) return ORX_E_TRUNC;
        uint8_t vb[48], hb[48];
        float llrht[96], llrvht[96];
        d.fftDemod(sig + SYM_SHIFT, d.fo1);
        d.fftDemod(sig + SYM_SHIFT + 80, d.fo2);
        nlsigDemod(d.fo1, d.fo2, d.HL, llrht, llrvht);
        sigViterbi(llrvht, vb, 48);
        if (checkVhtA(vb)) {                                  // DEMOD_S_VHT :150-178
            parseVhtA(vb, &d.m);
            pos = 160;
            int need = 80 + d.m.nLTF * 80 + 80;
            if (nsig - pos < need) return ORX_E_TRUNC;
            uint8_t sb[26];
            d.nonLegacyChanEstimate(sig + pos + 80);
            d.vhtSigBDemod(sig + pos + 80 + d.m.nLTF * 80, sb);
            parseVhtB(sb, &d.m);
            int nl = (llen * 8 + 22 + 23) / 24;
            bool ok = d.m.len >= 0 && d.m.len <= 4095 && d.m.nSS <= 2 &&
                      (nl * 80) >= (d.m.nSym * d.m.nSymSamp + 160 + 80 + d.m.nLTF * 80 + 80);
            pos += need;
            if (!ok) return ORX_E_FORMAT;
            trellis = d.m.nSym * d.m.nDBPS;
            d.pilot[0] = 1.f; d.pilot[1] = 1.f; d.pilot[2] = 1.f; d.pilot[3] = -1.f;   // PILOT_VHT
            d.pilotP = 4;
        } else {
            sigViterbi(llrht, hb, 48);
            if (checkHt(hb)) {                                // DEMOD_S_HT :180-205
                parseHt(hb, &d.m);
                pos = 160;
                int need = 80 + d.m.nLTF * 80;
                if (nsig - pos < need) return ORX_E_TRUNC;
                d.nonLegacyChanEstimate(sig + pos + 80);
                int nl = (llen * 8 + 22 + 23) / 24;
                bool ok = d.m.len > 0 && d.m.len <= 4095 && d.m.nSS <= 2 &&
                          (nl * 80) >= (d.m.nSym * d.m.nSymSamp + 160 + 80 + d.m.nLTF * 80);
                pos += need;
                if (!ok) return ORX_E_FORMAT;
                trellis = d.m.len * 8 + 22;
                d.pilot[0] = 1.f; d.pilot[1] = 1.f; d.pilot[2] = 1.f; d.pilot[3] = -1.f;   // PILOT_HT_1
                d.pilotP = 3;
            } else {
                legacy = true;
            }
        }
    }
    if (legacy) {                                             // DEMOD_S_LEGACY :207-219
        parseL(lmcs, llen, &d.m);
        trellis = d.m.len * 8 + 22;
        d.pilot[0] = 1.f; d.pilot[1] = 1.f; d.pilot[2] = 1.f; d.pilot[3] = -1.f;       // PILOT_L
        d.pilotP = 1;
    }
    // DEMOD_S_WRTAG :221-277
    f->format = d.m.format; f->mcs = d.m.mcs; f->len = d.m.len; f->cr = d.m.cr; f->ampdu = d.m.ampdu;
    f->nss = d.m.nSS; f->nsym = d.m.nSym; f->nsymsamp = d.m.nSymSamp; f->ncbps = d.m.nCBPS; f->ndbps = d.m.nDBPS;
    f->trellis = trellis; f->total = d.m.nSym * d.m.nCBPS; f->data_off = pos;
    f->sssnr0 = (d.m.format == F_VHT) ? d.sssnr : 0.f; f->sssnr1 = 0.f;
    if (d.m.nSym == 0) {                                      // :238-249, :264-270: tag mu2x1chan, 1024 items of no content
        f->total = 1024;
        // the tag carries d_mu2x1Chan, which only the sounding branch of nonLegacyChanEstimate fills (:391-394); for a
        // one-stream NDP the reference would publish whatever the member held before -- modelled as zeros
        if (llrOut && llrCap >= 256) {
            if (d.m.nSS == 1 && d.m.nLTF == 1) memset(llrOut, 0, 256 * sizeof(float));
            else memcpy(llrOut, d.mu2x1, 256 * sizeof(float));
        }
        return ORX_E_NDP;
    }
    if (f->total > llrCap) return ORX_E_TRUNC;
    // DEMOD_S_DEMOD :279-314 ; "(o1 + nSymSamp) < d_nProc" needs one spare input sample
    float inted[416];
    for (int sidx = 0; sidx < d.m.nSym; sidx++) {
        int o1 = pos + sidx * d.m.nSymSamp;
        if (o1 + d.m.nSymSamp > nsig) return ORX_E_TRUNC;
        float* out = llrOut + (size_t)sidx * d.m.nCBPS;
        if (d.m.format == F_L) {
            d.legacyChanUpdate(sig + o1);
            qamToLlr(d.qam, inted, d.m.mod, d.m.nSD);
            const int* map = T.deintL[nbIndexL(d.m.nBPSCS)];                 // procSymDeintL2 c8p.cc:2150-2192
            for (int i = 0; i < d.m.nCBPS; i++) out[map[i]] = inted[i];
        } else {
            d.nonLegacyChanUpdate(sig + o1);
            qamToLlr(d.qam, inted, d.m.mod, d.m.nSD);
            int bi = nbIndexNL(d.m.nBPSCS);                                   // procSymDeintNL2SS1 c8p.cc:2238-2287
            if (bi >= 0) { const int* map = T.deintNL[0][bi]; for (int i = 0; i < d.m.nCBPSS; i++) out[map[i]] = inted[i]; }
        }
    }
    return ORX_OK;
}

// ------------------------------------------------------------------------------------------------
// demod2: 2x2 SU-MIMO demod (ref: lib/demod2_impl.cc:58-348 state machine, :350-806 helpers).
// sig1/sig2 = the two CFO-compensated streams signal2 copies out (lib/signal2_impl.cc:164-192).
// ------------------------------------------------------------------------------------------------
struct Demod2 {
    Mod m;
    cf HL[64], H[64][4], HI[64][4], sig1[64], sig2[64], qam[2][52], f1[64], f2[64], f12[64], f22[64];
    cf pnl[4], pnl2[4];
    float pilot[4], pilot2[4];
    int pilotP;
    float sssnr0, sssnr1;

    static bool nlNull(int i) { return i == 0 || (i >= 29 && i <= 35); }
    static bool lNull(int i) { return i == 0 || (i >= 27 && i <= 37); }
    static bool isPilot(int i) { return i == 7 || i == 21 || i == 43 || i == 57; }
    static void shift(float* p) { float t = p[0]; p[0] = p[1]; p[1] = p[2]; p[2] = p[3]; p[3] = t; }

    void zf(int i, const cf& a1, const cf& a2, cf& s1, cf& s2)        // (H^H H)^-1 H^H y, :498-501
    {
        cf t1 = a1 * std::conj(H[i][0]) + a2 * std::conj(H[i][1]);
        cf t2 = a1 * std::conj(H[i][2]) + a2 * std::conj(H[i][3]);
        s1 = t1 * HI[i][0] + t2 * HI[i][2];
        s2 = t1 * HI[i][1] + t2 * HI[i][3];
    }

    void chanEstimate(const cf* s1, const cf* s2)                     // :350-469
    {
        if (m.nSS == 1) {
            if (m.nLTF == 1) {
                fft64(s1 + SYM_SHIFT, f1);
                for (int i = 0; i < 64; i++) if (!nlNull(i)) H[i][0] = f1[i] / T.ltfNL[i];
            }
        } else if (m.nSS == 2) {
            fft64(s1 + SYM_SHIFT, f1); fft64(s2 + SYM_SHIFT, f2);
            fft64(s1 + SYM_SHIFT + 80, f12); fft64(s2 + SYM_SHIFT + 80, f22);
            for (int i = 0; i < 64; i++) {
                if (nlNull(i)) continue;
                float l2 = T.ltfNL[i] * 0.5f;                         // LTF_NL_28_F_FLOAT2 (c8p.cc:142-158) = LTF/2, exact
                H[i][0] = (f1[i] - f12[i]) * l2; H[i][1] = (f2[i] - f22[i]) * l2;
                H[i][2] = (f1[i] + f12[i]) * l2; H[i][3] = (f2[i] + f22[i]) * l2;
            }
            if (m.format == F_VHT) {                                  // pilot tones interpolated :391-409
                const int pb[4] = { 7, 21, 43, 57 };
                for (int q = 0; q < 4; q++) for (int k = 0; k < 4; k++) H[pb[q]][k] = (H[pb[q] - 1][k] + H[pb[q] + 1][k]) / 2.0f;
            }
            for (int i = 0; i < 64; i++) {
                if (nlNull(i)) continue;
                cf a = H[i][0] * std::conj(H[i][0]) + H[i][1] * std::conj(H[i][1]);
                cf b = H[i][0] * std::conj(H[i][2]) + H[i][1] * std::conj(H[i][3]);
                cf c = H[i][2] * std::conj(H[i][0]) + H[i][3] * std::conj(H[i][1]);
                cf d = H[i][2] * std::conj(H[i][2]) + H[i][3] * std::conj(H[i][3]);
                cf inv = 1.0f / (a * d - b * c);
                HI[i][0] = inv * d; HI[i][1] = -inv * b; HI[i][2] = -inv * c; HI[i][3] = inv * a;
            }
            const int pb[4] = { 7, 21, 43, 57 }, slot[4] = { 2, 3, 0, 1 };
            for (int q = 0; q < 4; q++) {                             // pilot references from the first LTF :432-463
                cf t1, t2;
                zf(pb[q], f1[pb[q]], f2[pb[q]], t1, t2);
                if (q == 3) { t1 = -t1; t2 = -t2; }
                pnl[slot[q]] = std::conj(t1); pnl2[slot[q]] = std::conj(t2);
            }
        }
    }

    cf pilotSum2(const float* pa, const float* pb_)                   // 8-term sum of htChanUpdate / vhtChanUpdate
    {
        float P = T.pilotP[pilotP];
        return std::conj(sig1[7] * pa[2] * P * pnl[2] + sig1[21] * pa[3] * P * pnl[3] + sig1[43] * pa[0] * P * pnl[0] + sig1[57] * pa[1] * P * pnl[1] +
                         sig2[7] * pb_[2] * P * pnl2[2] + sig2[21] * pb_[3] * P * pnl2[3] + sig2[43] * pb_[0] * P * pnl2[0] + sig2[57] * pb_[1] * P * pnl2[1]);
    }

    void symUpdate(const cf* s1, const cf* s2)                        // htChanUpdate :471-551 / vhtChanUpdate :553-630
    {
        if (m.nSS == 1) {
            fft64(s1 + SYM_SHIFT, f1);
            for (int i = 0; i < 64; i++) if (!nlNull(i)) sig1[i] = f1[i] / H[i][0];
            float P = T.pilotP[pilotP];
            cf ps = std::conj(sig1[7] * pilot[2] * P + sig1[21] * pilot[3] * P + sig1[43] * pilot[0] * P + sig1[57] * pilot[1] * P);
            shift(pilot);
            pilotP = (pilotP + 1) % 127;
            float pa = std::abs(ps);
            int j = 26;
            for (int i = 0; i < 64; i++) { if (nlNull(i) || isPilot(i)) continue; qam[0][j] = sig1[i] * ps / pa; if (++j >= 52) j = 0; }
        } else {
            fft64(s1 + SYM_SHIFT, f1); fft64(s2 + SYM_SHIFT, f2);
            for (int i = 0; i < 64; i++) if (!nlNull(i)) zf(i, f1[i], f2[i], sig1[i], sig2[i]);
            cf ps = (m.format == F_VHT) ? pilotSum2(pilot, pilot) : pilotSum2(pilot, pilot2);
            shift(pilot);
            if (m.format != F_VHT) shift(pilot2);
            pilotP = (pilotP + 1) % 127;
            float pa = std::abs(ps);
            int j = 26;
            for (int i = 0; i < 64; i++) {
                if (nlNull(i) || isPilot(i)) continue;
                qam[0][j] = sig1[i] * ps / pa; qam[1][j] = sig2[i] * ps / pa;
                if (++j >= 52) j = 0;
            }
        }
    }

    void legacyUpdate(const cf* s1)                                   // :760-786
    {
        fft64(s1 + SYM_SHIFT, f1);
        for (int i = 0; i < 64; i++) if (!lNull(i)) sig1[i] = f1[i] / HL[i];
        float P = T.pilotP[pilotP];
        cf ps = std::conj(sig1[7] * pilot[2] * P + sig1[21] * pilot[3] * P + sig1[43] * pilot[0] * P + sig1[57] * pilot[1] * P);
        pilotP = (pilotP + 1) % 127;
        float pa = std::abs(ps);
        int j = 24;
        for (int i = 0; i < 64; i++) { if (lNull(i) || isPilot(i)) continue; qam[0][j] = sig1[i] * ps / pa; if (++j >= 48) j = 0; }
    }

    void sigB(const cf* s1, const cf* s2, uint8_t* bits26)            // vhtSigBDemod :632-758
    {
        cf q0[52], q1[52];
        float inted[52], coded[52];
        if (m.nSS == 1) {
            fft64(s1 + SYM_SHIFT, f1);
            for (int i = 0; i < 64; i++) if (!nlNull(i)) sig1[i] = f1[i] / H[i][0];
            cf ps = std::conj(sig1[7] - sig1[21] + sig1[43] + sig1[57]);
            float pa = std::abs(ps);
            int j = 26;
            for (int i = 0; i < 64; i++) { if (nlNull(i) || isPilot(i)) continue; q0[j] = sig1[i] * ps / pa; inted[j] = q0[j].real(); if (++j >= 52) j = 0; }
        } else if (m.nSS == 2) {
            fft64(s1 + SYM_SHIFT, f1); fft64(s2 + SYM_SHIFT, f2);
            for (int i = 0; i < 64; i++) if (!nlNull(i)) zf(i, f1[i], f2[i], sig1[i], sig2[i]);
            cf ps = std::conj(sig1[7] * pnl[2] - sig1[21] * pnl[3] + sig1[43] * pnl[0] + sig1[57] * pnl[1] +
                              sig2[7] * pnl2[2] - sig2[21] * pnl2[3] + sig2[43] * pnl2[0] + sig2[57] * pnl2[1]);
            float pa = std::abs(ps);
            int j = 26;
            for (int i = 0; i < 64; i++) {
                if (nlNull(i) || isPilot(i)) continue;
                q0[j] = sig1[i] * ps / pa; q1[j] = sig2[i] * ps / pa;
                inted[j] = (q0[j].real() + q1[j].real()) / 2.0f;
                if (++j >= 52) j = 0;
            }
        } else { memset(bits26, 0, 26); return; }
        for (int i = 0; i < 52; i++) coded[T.deintNL[0][0][i]] = inted[i];
        sigViterbi(coded, bits26, 26);
        uint8_t enc[52], intl[52];
        bcc(bits26, enc, 26);
        for (int i = 0; i < 52; i++) intl[i] = enc[T.deintNL[0][0][i]];
        double n0 = 0.0, n1 = 0.0;
        for (int i = 0; i < 52; i++) {
            cf ref = intl[i] ? cf(1.0f, 0.0f) : cf(-1.0f, 0.0f);
            q0[i] -= ref;
            n0 += (double)(q0[i].real() * q0[i].real() + q0[i].imag() * q0[i].imag());
            if (m.nSS == 2) { q1[i] -= ref; n1 += (double)(q1[i].real() * q1[i].real() + q1[i].imag() * q1[i].imag()); }
        }
        sssnr0 = (float)(log10(52.0 / n0) * 10.0);
        if (m.nSS == 2) sssnr1 = (float)(log10(52.0 / n1) * 10.0);
    }
};

int demodFrame2(const cf* sig1, const cf* sig2, int nsig, int lmcs, int llen, const cf* hl, orx_frame* f, float* llrOut, int llrCap)
{
    static thread_local Demod2 d;
    memset(&d.m, 0, sizeof(d.m));
    for (int i = 0; i < 64; i++) { d.HL[i] = hl[i]; d.sig1[i] = d.sig2[i] = cf(0.f, 0.f); for (int k = 0; k < 4; k++) d.H[i][k] = d.HI[i][k] = cf(0.f, 0.f); }
    d.sssnr0 = d.sssnr1 = 0.f;
    int pos = 0, trellis = 0;
    bool legacy = lmcs > 0;
    if (!legacy) {                                                    // DEMOD_S_FORMAT :104-147 (antenna 0 only)
        if (nsig < 160) return ORX_E_TRUNC;
        uint8_t vb[48], hb[48];
        float llrht[96], llrvht[96];
        fft64(sig1 + SYM_SHIFT, d.f1); fft64(sig1 + SYM_SHIFT + 80, d.f2);
        nlsigDemod(d.f1, d.f2, d.HL, llrht, llrvht);
        sigViterbi(llrvht, vb, 48);
        if (checkVhtA(vb)) {                                          // DEMOD_S_VHT :149-178
            parseVhtA(vb, &d.m);
            pos = 160;
            int need = 80 + d.m.nLTF * 80 + 80;
            if (nsig - pos < need) return ORX_E_TRUNC;
            uint8_t sb[26];
            d.chanEstimate(sig1 + pos + 80, sig2 + pos + 80);
            d.sigB(sig1 + pos + 80 + d.m.nLTF * 80, sig2 + pos + 80 + d.m.nLTF * 80, sb);
            parseVhtB(sb, &d.m);
            int nl = (llen * 8 + 22 + 23) / 24;
            bool ok = d.m.len > 0 && d.m.len <= 4095 && d.m.nSS <= 2 && (nl * 80) >= (d.m.nSym * d.m.nSymSamp + 160 + 80 + d.m.nLTF * 80 + 80);
            pos += need;
            if (!ok) return ORX_E_FORMAT;
            trellis = d.m.nSym * d.m.nDBPS;
            d.pilot[0] = 1.f; d.pilot[1] = 1.f; d.pilot[2] = 1.f; d.pilot[3] = -1.f;                 // PILOT_VHT
            d.pilotP = 4;
        } else {
            sigViterbi(llrht, hb, 48);
            if (checkHt(hb)) {                                        // DEMOD_S_HT :180-216
                parseHt(hb, &d.m);
                pos = 160;
                int need = 80 + d.m.nLTF * 80;
                if (nsig - pos < need) return ORX_E_TRUNC;
                d.chanEstimate(sig1 + pos + 80, sig2 + pos + 80);
                int nl = (llen * 8 + 22 + 23) / 24;
                bool ok = d.m.len > 0 && d.m.len <= 4095 && d.m.nSS <= 2 && (nl * 80) >= (d.m.nSym * d.m.nSymSamp + 160 + 80 + d.m.nLTF * 80);
                pos += need;
                if (!ok) return ORX_E_FORMAT;
                trellis = d.m.len * 8 + 22;
                if (d.m.nSS == 1) { d.pilot[0] = 1.f; d.pilot[1] = 1.f; d.pilot[2] = 1.f; d.pilot[3] = -1.f; }     // PILOT_HT_1
                else { d.pilot[0] = 1.f; d.pilot[1] = 1.f; d.pilot[2] = -1.f; d.pilot[3] = -1.f; }                 // PILOT_HT_2_1
                d.pilot2[0] = 1.f; d.pilot2[1] = -1.f; d.pilot2[2] = -1.f; d.pilot2[3] = 1.f;                      // PILOT_HT_2_2
                d.pilotP = 3;
            } else legacy = true;
        }
    }
    if (legacy) {                                                     // DEMOD_S_LEGACY :218-230
        parseL(lmcs, llen, &d.m);
        trellis = d.m.len * 8 + 22;
        d.pilot[0] = 1.f; d.pilot[1] = 1.f; d.pilot[2] = 1.f; d.pilot[3] = -1.f;
        d.pilotP = 1;
    }
    f->format = d.m.format; f->mcs = d.m.mcs; f->len = d.m.len; f->cr = d.m.cr; f->ampdu = d.m.ampdu;
    f->nss = d.m.nSS; f->nsym = d.m.nSym; f->nsymsamp = d.m.nSymSamp; f->ncbps = d.m.nCBPS; f->ndbps = d.m.nDBPS;
    f->trellis = trellis; f->total = d.m.nSym * d.m.nCBPS; f->data_off = pos;
    f->sssnr0 = (d.m.format == F_VHT) ? d.sssnr0 : 0.f;
    f->sssnr1 = (d.m.format == F_VHT && d.m.nSS == 2) ? d.sssnr1 : 0.f;
    if (f->total > llrCap) return ORX_E_TRUNC;
    float inted[2][416], spasd[2][416];
    for (int sidx = 0; sidx < d.m.nSym; sidx++) {                     // DEMOD_S_DEMOD :279-330
        int o1 = pos + sidx * d.m.nSymSamp;
        if (o1 + d.m.nSymSamp > nsig) return ORX_E_TRUNC;
        float* out = llrOut + (size_t)sidx * d.m.nCBPS;
        if (d.m.format == F_L) {
            d.legacyUpdate(sig1 + o1);
            qamToLlr(d.qam[0], inted[0], d.m.mod, d.m.nSD);
            const int* map = T.deintL[nbIndexL(d.m.nBPSCS)];
            for (int i = 0; i < d.m.nCBPS; i++) out[map[i]] = inted[0][i];
        } else {
            d.symUpdate(sig1 + o1, sig2 + o1);
            int bi = nbIndexNL(d.m.nBPSCS);
            if (bi < 0) continue;
            if (d.m.nSS == 1) {
                qamToLlr(d.qam[0], inted[0], d.m.mod, d.m.nSD);
                const int* map = T.deintNL[0][bi];
                for (int i = 0; i < d.m.nCBPSS; i++) out[map[i]] = inted[0][i];
            } else {
                qamToLlr(d.qam[0], inted[0], d.m.mod, d.m.nSD);
                qamToLlr(d.qam[1], inted[1], d.m.mod, d.m.nSD);
                for (int i = 0; i < d.m.nCBPSS; i++) { spasd[0][T.deintNL[0][bi][i]] = inted[0][i]; spasd[1][T.deintNL[1][bi][i]] = inted[1][i]; }
                int s = std::max(d.m.nBPSCS / 2, 1);                  // procSymDepasNL c8p.cc:2442-2451
                for (int i = 0; i < d.m.nCBPSS / s; i++) {
                    memcpy(&out[i * 2 * s], &spasd[0][i * s], sizeof(float) * s);
                    memcpy(&out[(i * 2 + 1) * s], &spasd[1][i * s], sizeof(float) * s);
                }
            }
        }
    }
    return ORX_OK;
}

// ------------------------------------------------------------------------------------------------
// decode (ref: lib/decode_impl.cc:164-203 vstb_init, :205-281 vstb_update, :282-302 vstb_end,
// :304-323 descramble, :325-520 packetAssemble)
// ------------------------------------------------------------------------------------------------
const int PUNC[4][10] = { { 1, 1 }, { 1, 1, 1, 0 }, { 1, 1, 1, 0, 0, 1 }, { 1, 1, 1, 0, 0, 1, 1, 0, 0, 1 } };   // c8p.cc:1857-1860
const int PUNC_LEN[4] = { 2, 4, 6, 10 };

void viterbi(const float* llr, int cr, int trellis, uint8_t* bits)
{
    std::vector<uint8_t> his((size_t)(trellis + 1) * 64, 0);
    float m0[64], m1[64];
    for (int i = 0; i < 64; i++) m0[i] = -1000000000000000.0f;
    m0[0] = 0;
    float *pre = m0, *cur = m1;
    const int* pn = PUNC[cr];
    int plen = PUNC_LEN[cr], p = 0, used = 0;
    for (int t = 0; t < trellis; t++) {
        float t0 = 0.0f, t1 = 0.0f;
        if (pn[p]) t0 = llr[used++];                          // :212-229 erased positions are 0.0f
        if (pn[p + 1]) t1 = llr[used++];
        acsStep(pre, cur, &his[(size_t)(t + 1) * 64], t0, t1);
        std::swap(pre, cur);
        p += 2; if (p >= plen) p = 0;
    }
    traceback(his.data(), trellis, bits);
}

void descramble(const uint8_t* sb, int trellis, uint8_t* ub)   // :304-323
{
    int state = 0;
    for (int i = 0; i < 7 && i < trellis; i++) if (sb[i]) state |= 1 << (6 - i);
    for (int i = 0; i < 7 && i < trellis; i++) ub[i] = 0;
    for (int i = 7; i < trellis; i++) {
        int fb = ((state >> 6) & 1) ^ ((state >> 3) & 1);
        ub[i] = fb ^ (sb[i] & 1);
        state = ((state << 1) & 0x7e) | fb;
    }
}

uint32_t crc32(const uint8_t* p, int n)    // boost::crc_32_type (decode_impl.h:84): reflected 0x04C11DB7, init/xorout ~0
{
    uint32_t c = 0xFFFFFFFFu;
    for (int i = 0; i < n; i++) c = T.crc32tab[(c ^ p[i]) & 0xff] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

// packetAssemble :325-520.  Appends records [format][len lo][len hi][MPDU][mcs]; returns count.
int packetAssemble(const uint8_t* ub, int trellis, int format, int len, int mcs, int ampdu, uint8_t* out, int cap, int* used)
{
    int npdu = 0, w = 0;
    std::vector<uint8_t> pkt(16384 + 8);
    if (format == F_VHT) {
        int procd = 16;
        if (procd < trellis) {
            const uint8_t* bp = ub + 16;
            int eof, tl = 0;                                   // tl is NOT reset per subframe (:336)
            while (true) {
                procd += 32;
                if (procd > trellis) break;
                eof = bp[0];
                tl |= ((int)bp[2]) << 12;
                tl |= ((int)bp[3]) << 13;
                for (int i = 0; i < 12; i++) tl |= ((int)bp[4 + i]) << i;
                int padded = (tl / 4 + ((tl % 4) != 0)) * 4 * 8;
                procd += padded;
                if (procd > trellis) break;
                bp += 32;
                pkt[0] = (uint8_t)format; pkt[1] = tl % 256; pkt[2] = tl / 256;
                for (int i = 0; i < tl; i++) { uint8_t v = 0; for (int j = 0; j < 8; j++) v |= bp[i * 8 + j] << j; pkt[i + 3] = v; }
                bp += padded;
                if (crc32(&pkt[3], tl) == 558161692u) {
                    pkt[tl + 3] = (uint8_t)mcs;
                    tl += 4;                                   // :415 (carries into the next subframe)
                    if (w + tl <= cap) { memcpy(out + w, pkt.data(), tl); w += tl; npdu++; }
                }
                if (eof) break;
            }
        }
    } else if (!ampdu) {
        const uint8_t* bp = ub + 16;
        pkt[0] = (uint8_t)format; pkt[1] = len % 256; pkt[2] = len / 256;
        for (int i = 0; i < len; i++) { uint8_t v = 0; for (int j = 0; j < 8; j++) v |= bp[i * 8 + j] << j; pkt[i + 3] = v; }
        if (crc32(&pkt[3], len) == 558161692u) {
            pkt[3 + len] = (uint8_t)mcs;
            if (w + len + 4 <= cap) { memcpy(out + w, pkt.data(), len + 4); w += len + 4; npdu++; }
        }
    }
    *used = w;
    return npdu;
}

int decodeFrame(const float* llr, const orx_frame* f, uint8_t* pdu, int cap, int* used, uint8_t* sbOpt)
{
    *used = 0;
    if (f->len > DECODE_B_MAX || f->trellis > DECODE_T_MAX) return -1;    // :93-97
    if (f->trellis <= 0) return 0;
    std::vector<uint8_t> sb(f->trellis), ub(f->trellis);
    viterbi(llr, f->cr, f->trellis, sb.data());
    if (sbOpt) memcpy(sbOpt, sb.data(), f->trellis);
    descramble(sb.data(), f->trellis, ub.data());
    return packetAssemble(ub.data(), f->trellis, f->format, f->len, f->mcs, f->ampdu, pdu, cap, used);
}

// ------------------------------------------------------------------------------------------------
// whole chain on one item: the block state machines evaluated over one capture segment from reset
// (sync IDLE/SYNC: lib/sync_impl.cc:73-147; signal S_TRIGGER/S_DEMOD/S_COPY/S_PAD:
// lib/signal_impl.cc:75-202).  Stalls at the end of the segment (not enough samples left for a
// state to proceed) end the item, as the flowgraph ends when the file source runs dry.
// ------------------------------------------------------------------------------------------------
struct SyncEv { int trig, idx; float rad, snr, rssi; };

int rxItem(const cf* x, int n, int item, int maxFrames, orx_frame* frames, float* llr, int64_t llrCap, int64_t* llrUsed,
           uint8_t* pdu, int64_t pduCap, int64_t* pduUsed, std::vector<float>* llrScratch, const cf* x2 = nullptr)
{
    std::vector<float> preac(n);
    std::vector<cf> preconj(n);
    std::vector<uint8_t> trig(n);
    presiso(x, n, preac.data(), preconj.data());
    TrigState ts; memset(&ts, 0, sizeof(ts));
    trigger(preac.data(), n, trig.data(), ts);

    std::vector<SyncEv> evs;
    int nTrig = 0;
    {
        cf latch(0.f, 0.f);
        int i = 0;
        while (i < n) {
            if (trig[i] & 0x01) {
                nTrig++;
                if (n - i < SYNC_BUF) break;                   // sync_impl.cc:94 never satisfied -> stall
                SyncOut so = syncAt(x + i, latch, nullptr);
                if (so.ok) { SyncEv e; e.trig = i; e.idx = i + so.mIndex; e.rad = so.rad; e.snr = so.snr; e.rssi = so.rssi; evs.push_back(e); }
                i += SYNC_RES;
                continue;
            } else if (trig[i] & 0x02) {
                latch = preconj[i];
            }
            i++;
        }
    }

    int nf = 0, nLsigFail = 0;
    int pos = 0;
    std::vector<cf> rot, rot2;
    for (size_t e = 0; e < evs.size() && nf < maxFrames; e++) {
        const SyncEv& ev = evs[e];
        if (ev.idx < pos) continue;                            // swallowed by S_COPY / skipped 80
        if (n - ev.idx < 224) break;                           // signal_impl.cc:110 stall
        cf h[64];
        float llr48[48];
        uint8_t bits[24];
        int mcs = 0, len = 0, nsamp = 0;
        if (!signalAt(x + ev.idx, ev.rad, h, llr48, bits, &mcs, &len, &nsamp)) { nLsigFail++; pos = ev.idx + 80; continue; }
        orx_frame* f = &frames[nf++];
        memset(f, 0, sizeof(*f));
        f->item = item; f->trig_idx = ev.trig; f->sync_idx = ev.idx; f->rad = ev.rad; f->snr = ev.snr; f->rssi = ev.rssi;
        f->cfo_hz = ev.rad * 3183098.8618379068f;              // signal_impl.cc:135
        f->l_mcs = mcs; f->l_len = len; f->nsamp = nsamp;
        f->llr_off = *llrUsed; f->pdu_off = *pduUsed;
        int start = ev.idx + 224;
        pos = start + nsamp;
        if (pos > n) { f->status = ORX_E_TRUNC; break; }       // S_COPY can never finish
        // S_COPY :164-192 (+ S_PAD: 320 unwritten samples; modelled as zeros, never read by a passing frame)
        rot.assign((size_t)nsamp + 320, cf(0.f, 0.f));
        for (int k = 0; k < nsamp; k++) {
            float ph = (float)(k + 224) * ev.rad;
            rot[k] = x[start + k] * cf(cosf(ph), sinf(ph));
        }
        int64_t room = llrCap - *llrUsed;
        float* lo = llr ? llr + *llrUsed : nullptr;
        int capI = (int)std::min<int64_t>(room, 1 << 30);
        if (!llr) { llrScratch->resize(1366 * 832 + 1024); lo = llrScratch->data(); capI = (int)llrScratch->size(); }
        if (x2) std::fill(lo, lo + std::min<int64_t>(capI, (int64_t)(nsamp / 72 + 2) * 832), 0.0f);
        if (x2) {                                              // signal2: same rotation on antenna 1 (lib/signal2_impl.cc:177,190)
            rot2.assign((size_t)nsamp + 320, cf(0.f, 0.f));
            for (int k = 0; k < nsamp; k++) {
                float ph = (float)(k + 224) * ev.rad;
                rot2[k] = x2[start + k] * cf(cosf(ph), sinf(ph));
            }
            f->status = demodFrame2(rot.data(), rot2.data(), nsamp + 320, mcs, len, h, f, lo, capI);
        } else
            f->status = demodFrame(rot.data(), nsamp + 320, mcs, len, h, f, lo, capI);
        if (f->status == ORX_E_NDP && !x2 && pduCap - *pduUsed >= 1027 && capI >= 256) {
            // decode_impl.cc:100-121: v_trellis == 0 -> channel report [C8P_F_VHT_CHAN = 20][len lo][len hi][128 x (re, im) float]
            uint8_t* p = pdu + *pduUsed;
            p[0] = 20; p[1] = 1024 % 256; p[2] = 1024 / 256;
            memcpy(p + 3, lo, 1024);
            f->npdu = 1; f->pdu_bytes = 1027;
            *pduUsed += 1027;
            continue;
        }
        if (f->status != ORX_OK) continue;
        if (llr) *llrUsed += f->total;
        int used = 0;
        int np = decodeFrame(lo, f, pdu + *pduUsed, (int)std::min<int64_t>(pduCap - *pduUsed, 1 << 30), &used, nullptr);
        if (np < 0) { f->status = ORX_E_DECODE_RANGE; continue; }
        f->npdu = np; f->pdu_bytes = used;
        *pduUsed += used;
    }
    if (nf == 0) {
        orx_frame* f = &frames[0];
        memset(f, 0, sizeof(*f));
        f->item = item;
        f->status = nTrig == 0 ? ORX_E_NO_TRIGGER : evs.empty() ? ORX_E_SYNC : (nLsigFail ? ORX_E_LSIG : ORX_E_TRUNC);
        f->llr_off = *llrUsed; f->pdu_off = *pduUsed;
        nf = 1;
    }
    return nf;
}

} // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int orx_deint_map(int nonlegacy, int nbpscs, int iss, int* out)
{
    if (!nonlegacy) { int b = nbIndexL(nbpscs); if (b < 0) return -1; memcpy(out, T.deintL[b], sizeof(int) * 48 * nbpscs); return 48 * nbpscs; }
    int b = nbIndexNL(nbpscs); if (b < 0 || iss < 1 || iss > 2) return -1;
    memcpy(out, T.deintNL[iss - 1][b], sizeof(int) * 52 * nbpscs);
    return 52 * nbpscs;
}
void orx_pilot_polarity(float* p) { memcpy(p, T.pilotP, sizeof(T.pilotP)); }
void orx_ltf(int kind, float* f) { memcpy(f, kind == 0 ? T.ltfL : kind == 1 ? T.ltfNL : T.ltfNL22, sizeof(float) * 64); }
void orx_trellis_tables(int* nx, int* op) { for (int i = 0; i < 128; i++) { nx[i] = T.nxt[i / 2][i % 2]; op[i] = T.outp[i / 2][i % 2]; } }
void orx_lsig_demap(int* o) { memcpy(o, T.lsigDemap, sizeof(T.lsigDemap)); }
void orx_set_mupos(int p) { g_mupos = p; }

void orx_fft64(const float* in, float* out) { fft64(reinterpret_cast<const cf*>(in), reinterpret_cast<cf*>(out)); }
void orx_presiso(const float* iq, int n, float* preac, float* preconj) { presiso(reinterpret_cast<const cf*>(iq), n, preac, reinterpret_cast<cf*>(preconj)); }
void orx_trigger(const float* preac, int n, uint8_t* out, int32_t* st)
{
    TrigState s; s.nPlateau = st[0]; s.fPlateau = st[1]; s.fPlateauEnd = st[2]; s.countDown = st[3]; memcpy(&s.conjAc, &st[4], 4);
    trigger(preac, n, out, s);
    st[0] = s.nPlateau; st[1] = s.fPlateau; st[2] = s.fPlateauEnd; st[3] = s.countDown; memcpy(&st[4], &s.conjAc, 4);
}
int orx_sync(const float* iq, float cre, float cim, int* mIndex, float* rad, float* snr, float* rssi, float* ac111)
{
    SyncOut o = syncAt(reinterpret_cast<const cf*>(iq), cf(cre, cim), ac111);
    *mIndex = o.mIndex; *rad = o.rad; *snr = o.snr; *rssi = o.rssi;
    return o.ok;
}
int orx_signal(const float* iq, float rad, float* h, float* llr48, uint8_t* bits24, int* mcs, int* len, int* nsamp)
{
    return signalAt(reinterpret_cast<const cf*>(iq), rad, reinterpret_cast<cf*>(h), llr48, bits24, mcs, len, nsamp);
}
void orx_lsig_demod(const float* s1, const float* s2, const float* sig, float* h, float* llr48)
{
    cf hh[64]; for (int i = 0; i < 64; i++) hh[i] = cf(0.f, 0.f);
    lsigDemod(reinterpret_cast<const cf*>(s1), reinterpret_cast<const cf*>(s2), reinterpret_cast<const cf*>(sig), hh, llr48);
    memcpy(h, hh, sizeof(hh));
}
void orx_nlsig_demod(const float* s1, const float* s2, const float* h, float* llrht, float* llrvht)
{
    nlsigDemod(reinterpret_cast<const cf*>(s1), reinterpret_cast<const cf*>(s2), reinterpret_cast<const cf*>(h), llrht, llrvht);
}
void orx_sig_viterbi(const float* llr, uint8_t* bits, int trellis) { sigViterbi(llr, bits, trellis); }
int orx_crc8_check(const uint8_t* bits, int len, const uint8_t* crc) { return crc8Check(bits, len, crc) ? 1 : 0; }
int orx_check_legacy(const uint8_t* b, int* mcs, int* len, int* ndbps) { return lsigCheck(b, mcs, len, ndbps) ? 1 : 0; }
int orx_check_ht(const uint8_t* b) { return checkHt(b) ? 1 : 0; }
int orx_check_vhta(const uint8_t* b) { return checkVhtA(b) ? 1 : 0; }
// parsers export the 17 SU ints of Mod: format,sumu,ampdu,nSym,nSymSamp,nSD,nSP,nSS,nLTF,mcs,len,mod,cr,nBPSCS,nDBPS,nCBPS,nCBPSS
void orx_parse_l(int mcs, int len, int* m17) { Mod m; memset(&m, 0, sizeof(m)); parseL(mcs, len, &m); memcpy(m17, &m, sizeof(m)); }
void orx_parse_ht(const uint8_t* b, int* m17) { Mod m; memset(&m, 0, sizeof(m)); parseHt(b, &m); memcpy(m17, &m, sizeof(m)); }
void orx_parse_vhta(const uint8_t* b, int* m17) { Mod m; memset(&m, 0, sizeof(m)); parseVhtA(b, &m); memcpy(m17, &m, sizeof(m)); }
void orx_parse_vhtb(const uint8_t* b, int* m17) { Mod m; memcpy(&m, m17, sizeof(m)); parseVhtB(b, &m); memcpy(m17, &m, sizeof(m)); }

int orx_demod(const float* sig, int nsig, int lmcs, int llen, const float* h, orx_frame* f, float* llrOut, int llrCap)
{
    return demodFrame(reinterpret_cast<const cf*>(sig), nsig, lmcs, llen, reinterpret_cast<const cf*>(h), f, llrOut, llrCap);
}
void orx_qam_to_llr(float* qam, float* llr, int mod, int nsd) { qamToLlr(reinterpret_cast<cf*>(qam), llr, mod, nsd); }
void orx_viterbi(const float* llr, int cr, int trellis, uint8_t* sb) { viterbi(llr, cr, trellis, sb); }
void orx_descramble(const uint8_t* sb, int trellis, uint8_t* b) { descramble(sb, trellis, b); }
int orx_decode(const float* llr, const orx_frame* f, uint8_t* pdu, int cap, int* used, uint8_t* sbOpt) { return decodeFrame(llr, f, pdu, cap, used, sbOpt); }
uint32_t orx_crc32(const uint8_t* p, int n) { return crc32(p, n); }

int orx_rx_item(const float* iq, int n, int item, int maxFrames, orx_frame* frames, float* llr, int64_t llrCap, int64_t* llrUsed,
                uint8_t* pdu, int64_t pduCap, int64_t* pduUsed)
{
    std::vector<float> scratch;
    return rxItem(reinterpret_cast<const cf*>(iq), n, item, maxFrames, frames, llr, llrCap, llrUsed, pdu, pduCap, pduUsed, &scratch);
}

int orx_rx_item2(const float* iq0, const float* iq1, int n, int item, int maxFrames, orx_frame* frames, float* llr, int64_t llrCap,
                 int64_t* llrUsed, uint8_t* pdu, int64_t pduCap, int64_t* pduUsed)
{
    std::vector<float> scratch;
    return rxItem(reinterpret_cast<const cf*>(iq0), n, item, maxFrames, frames, llr, llrCap, llrUsed, pdu, pduCap, pduUsed, &scratch,
                  reinterpret_cast<const cf*>(iq1));
}
int orx_demod2(const float* sig1, const float* sig2, int nsig, int lmcs, int llen, const float* h, orx_frame* f, float* llrOut, int llrCap)
{
    return demodFrame2(reinterpret_cast<const cf*>(sig1), reinterpret_cast<const cf*>(sig2), nsig, lmcs, llen, reinterpret_cast<const cf*>(h), f,
                       llrOut, llrCap);
}

int orx_rx_batch(const float* iq, const int64_t* offs, const int32_t* lens, int nitems, int nthreads, orx_frame* frames, uint8_t* pdu,
                 int64_t pduStride)
{
    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    std::atomic<int> next(0);
    auto worker = [&]() {
        std::vector<float> scratch;
        for (;;) {
            int i0 = next.fetch_add(4);
            if (i0 >= nitems) break;
            for (int i = i0; i < std::min(i0 + 4, nitems); i++) {
                int64_t lu = 0, pu = 0;
                rxItem(reinterpret_cast<const cf*>(iq) + offs[i], lens[i], i, 1, &frames[i], nullptr, 0, &lu,
                       pdu + (size_t)i * pduStride, pduStride, &pu, &scratch);
                frames[i].pdu_off = (int64_t)i * pduStride;
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; t++) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
    return nitems;
}

} // extern "C"
