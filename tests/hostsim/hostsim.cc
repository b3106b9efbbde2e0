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

// ---- the block state machines of gr-ieee80211_b200/csrc/blocks.h over a host backend ---------------------------------
// Per-frame arithmetic = the host build of phy_serial.cuh (the same routines the product's kernels run); the per-symbol
// demod and the Viterbi decode have no host build, so the backend returns zero soft bits and a placeholder PDU record
// [format][len lo][len hi][len x 0xA5][mcs]: this checks WHAT each scheduler call consumes, produces and tags, not bytes.
#include "../../gr-ieee80211_b200/csrc/blocks.h"

namespace {
struct HostBlk {
    int kind = 0, mupos = 0;
    TrigState ts;
    c8b_blocks::SyncState sy;
    c8b_blocks::SignalState sg;
    c8b_blocks::DemodState dm;
    c8b_blocks::DecodeState dc;

    int trigger(const float* in, int n, uint8_t* out) { for (int i = 0; i < n; i++) out[i] = trig_step(ts, in[i]); return 0; }
    int sync_at(const float* sig, const float conj[2], c8b_blocks::SyncRes* r)
    {
        const SyncOut o = c8b::sync_at((const cf*)sig, mk(conj[0], conj[1]));
        r->ok = o.ok; r->mIndex = o.mIndex; r->rad = o.rad; r->snr = o.snr; r->rssi = o.rssi;
        return 0;
    }
    int signal_at(const float* in, const float*, int, float rad, c8b_blocks::SignalRes* r, const float** rot0, const float** rot1)
    {
        *rot0 = *rot1 = nullptr;                                  // (this backend corrects the S_COPY samples in cfo_copy)
        cf h[64];
        r->mcs = r->len = r->nsamp = 0;
        r->ok = c8b::signal_at(lut(), (const cf*)in, rad, h, &r->mcs, &r->len, &r->nsamp);
        memcpy(r->chan, h, sizeof(h));
        return 0;
    }
    int cfo_copy(const float* in0, const float* in1, float* out0, float* out1, int n, int copied, float rad)
    {
        for (int i = 0; i < n; i++) {
            const cf w = cis(fmul((float)(copied + i + 224), rad));
            ((cf*)out0)[i] = cmul(((const cf*)in0)[i], w);
            if (in1) ((cf*)out1)[i] = cmul(((const cf*)in1)[i], w);
        }
        return 0;
    }
    // demod: header states on the host at submission (zero soft bits); collect hands the queue out in order
    std::vector<c8b_frame> mq;
    std::vector<float> msoft;
    int demod_submit(int nant, const float* iq0, const float* iq1, int, const c8b_frame* fin, const float* chan)
    {
        c8b_frame f = *fin;
        float hinv[128], w2[528];
        if (nant == 2) hs_header2(iq0, iq1, &f, chan, hinv, w2);
        else hs_header(iq0, &f, chan, mupos, hinv);
        mq.push_back(f);
        return 0;
    }
    int demod_collect(bool, c8b_frame* f, const float** soft, int* nsoft)
    {
        if (mq.empty()) return 0;
        *f = mq.front();
        mq.erase(mq.begin());
        msoft.assign((size_t)std::max(f->total, 1024), 0.f);
        *soft = msoft.data();
        *nsoft = (int)msoft.size();
        return 1;
    }
    // decode: the placeholder record is made at submission; collect hands the queue out in order
    std::vector<std::pair<c8b_frame, std::vector<uint8_t>>> dq;
    std::vector<uint8_t> cur;
    int decode_submit(const c8b_frame* fin, const float*, int)
    {
        c8b_frame f = *fin;
        std::vector<uint8_t> pdu((size_t)f.len + 4);
        pdu[0] = (uint8_t)f.format; pdu[1] = (uint8_t)(f.len & 255); pdu[2] = (uint8_t)(f.len >> 8);
        memset(pdu.data() + 3, 0xA5, (size_t)f.len);
        pdu[3 + f.len] = (uint8_t)f.mcs;
        f.npdu = 1; f.pdu_bytes = f.len + 4;
        dq.emplace_back(f, pdu);
        return 0;
    }
    int decode_collect(bool, c8b_frame* f, const uint8_t** pdu)
    {
        if (dq.empty()) return 0;
        *f = dq.front().first;
        cur = dq.front().second;
        dq.erase(dq.begin());
        *pdu = cur.data();
        return 1;
    }
};
}  // namespace

extern "C" {

void* hs_blk_create(int kind, int mupos) { HostBlk* b = new HostBlk; b->kind = kind; b->mupos = mupos; trig_reset(b->ts); return b; }
void hs_blk_destroy(void* b) { delete static_cast<HostBlk*>(b); }
int hs_blk_work(void* hb, int noutput, const int* ninput, const void* const* in, void* const* out, const c8b_tag* in_tags, int n_in_tags,
                int* consumed, int* produced, c8b_tag* out_tags, int out_tag_cap, int* n_out_tags, uint8_t* msg, int msg_cap, int* msg_bytes)
{
    HostBlk* b = static_cast<HostBlk*>(hb);
    c8b_blocks::WorkIO io;
    io.noutput = noutput; io.ninput = ninput; io.in = in; io.out = out; io.in_tags = in_tags; io.n_in_tags = n_in_tags;
    io.out_tags = out_tags; io.out_tag_cap = out_tag_cap; io.msg = msg; io.msg_cap = msg_cap;
    int rc;
    switch (b->kind) {
    case C8B_BLK_TRIGGER: rc = c8b_blocks::trigger_work(*b, io); break;
    case C8B_BLK_SYNC:    rc = c8b_blocks::sync_work(*b, b->sy, io); break;
    case C8B_BLK_SIGNAL:  rc = c8b_blocks::signal_work(*b, b->sg, 1, io); break;
    case C8B_BLK_SIGNAL2: rc = c8b_blocks::signal_work(*b, b->sg, 2, io); break;
    case C8B_BLK_DEMOD:   rc = c8b_blocks::demod_work(*b, b->dm, 1, io); break;
    case C8B_BLK_DEMOD2:  rc = c8b_blocks::demod_work(*b, b->dm, 2, io); break;
    default:              rc = c8b_blocks::decode_work(*b, b->dc, io); break;
    }
    *consumed = io.consumed; *produced = io.produced; *n_out_tags = io.n_out_tags; *msg_bytes = io.msg_bytes;
    return rc;
}

}
