// TEST INFRASTRUCTURE: compiles the product's per-frame routines (gr-ieee80211_b200/csrc/phy_serial.cuh,
// the code k_detect / k_header run one-thread-per-frame on the GPU) for the host, so the `-m "not gpu"`
// suite can check that host logic against the oracle.  Never loaded by the product.
#include "../../gr-ieee80211_b200/csrc/phy_serial.cuh"

#include <string.h>

using namespace c8b;

static c8b_lut g_lut;
static bool g_init = false;
static const c8b_lut* lut() { if (!g_init) { c8b_lut_build(&g_lut); g_init = true; } return &g_lut; }

struct RotSrc {
    const cf* x; float rad; int nsamp;
    cf operator()(int k) const { if (k >= nsamp) return mk(0.f, 0.f); return cmul(x[k], cis(fmul((float)(k + 224), rad))); }
};

extern "C" {

void hs_fft64(const float* in, float* out) { fft64(lut(), (const cf*)in, (cf*)out); }
void hs_conj_at(const float* iq, int i, float* out) { cf c = presiso_conj_at((const cf*)iq, i); out[0] = c.re; out[1] = c.im; }
void hs_trigger(const float* preac, int n, uint8_t* out) { TrigState s; trig_reset(s); for (int i = 0; i < n; i++) out[i] = trig_step(s, preac[i]); }
int hs_sync(const float* iq, float cre, float cim, int* mIndex, float* rad, float* snr, float* rssi)
{
    SyncOut o = sync_at((const cf*)iq, mk(cre, cim));
    *mIndex = o.mIndex; *rad = o.rad; *snr = o.snr; *rssi = o.rssi;
    return o.ok;
}
void hs_sig_viterbi(const float* llr, uint8_t* bits, int T) { sig_viterbi(lut(), llr, bits, T); }
int hs_crc8(const uint8_t* bits, int len, const uint8_t* crc) { return crc8_check(bits, len, crc) ? 1 : 0; }
void hs_detect(const float* iq, const float* preac, int n, int item, int maxf, c8b_frame* f, float* chan)
{
    memset(f, 0, sizeof(*f) * maxf);
    if (item < 0) {                                            // item < 0: scan with the threshold bitmap (as k_presiso + k_detect do)
        uint32_t* mask = new uint32_t[n / 32 + 2]();
        for (int i = 0; i < n; i++) if (preac[i] > 0.3f) mask[i >> 5] |= 1u << (i & 31);
        detect_item(lut(), (const cf*)iq, preac, n, -item - 1, maxf, f, (cf*)chan, mask);
        delete[] mask;
        return;
    }
    detect_item(lut(), (const cf*)iq, preac, n, item, maxf, f, (cf*)chan);
}
// window of a live stream: scan = {from, pos0, flush | safe, pos, nf, stalled} (c8b_scan, lut.h), bitmap scan like the GPU path
void hs_detect_scan(const float* iq, const float* preac, int n, int maxf, c8b_frame* f, float* chan, int32_t* scan, int bitmap)
{
    memset(f, 0, sizeof(*f) * maxf);
    uint32_t* mask = nullptr;
    if (bitmap) {
        mask = new uint32_t[n / 32 + 2]();
        for (int i = 0; i < n; i++) if (preac[i] > 0.3f) mask[i >> 5] |= 1u << (i & 31);
    }
    detect_item(lut(), (const cf*)iq, preac, n, 0, maxf, f, (cf*)chan, mask, reinterpret_cast<c8b_scan*>(scan));
    delete[] mask;
}
void hs_header(const float* iq_item, c8b_frame* f, const float* chan, int mupos, float* hinv)
{
    if (f->status != C8B_ST_OK) return;
    RotSrc rot; rot.x = (const cf*)iq_item + f->sync_idx + 224; rot.rad = f->rad; rot.nsamp = f->nsamp;
    f->status = demod_header(lut(), rot, f->nsamp, f->l_mcs, f->l_len, (const cf*)chan, mupos, f, (cf*)hinv);
}

void hs_header2(const float* iq0_item, const float* iq1_item, c8b_frame* f, const float* chan, float* hinv, float* w2)
{
    if (f->status != C8B_ST_OK) return;
    RotSrc r0; r0.x = (const cf*)iq0_item + f->sync_idx + 224; r0.rad = f->rad; r0.nsamp = f->nsamp;
    RotSrc r1 = r0; r1.x = (const cf*)iq1_item + f->sync_idx + 224;
    f->status = demod_header2(lut(), r0, r1, f->nsamp, f->l_mcs, f->l_len, (const cf*)chan, f, (cf*)hinv, (cf*)w2);
}

}
