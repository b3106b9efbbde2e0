// blocks.h -- the state machines of the seven receive blocks as the reference's general_work() runs them, one call of
// the GNU Radio scheduler at a time: which items a call consumes and produces, where it puts its tags, what it keeps
// between calls.  Host C++ only; every piece of arithmetic goes through an `Ops` backend:
//   * blocks.cu      -- the product: Ops = sm_100a kernels on the block's own c8b_ctx (C ABI c8b_blk_*)
//   * tests/hostsim  -- test infrastructure: Ops = the same per-frame routines compiled for the host
// Reference bodies followed here (file:line): lib/trigger_impl.cc:59-117, lib/sync_impl.cc:61-153,
// lib/signal_impl.cc:62-206, lib/signal2_impl.cc:63-212, lib/demod_impl.cc:59-342, lib/demod2_impl.cc:58-348,
// lib/decode_impl.cc:60-162.
//
// Deviation from the reference that a downstream block cannot see: demod / demod2 / decode gather a whole frame before
// they run their kernels (one launch per frame instead of one FFT per symbol per call), so the SAME items and tags leave
// the block, but later within the stream of scheduler calls; decode publishes a frame's MPDUs one call after its last soft
// bit arrived (or when called without input), while the GPU already works on it.  The 320 pad samples signal never writes
// (lib/signal_impl.cc:194-201) and the 1024 NDP floats demod never writes (lib/demod_impl.cc:251-257) are zeros here.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/c80211b200.h"

namespace c8b_blocks {

struct WorkIO {
    int noutput = 0;
    const int* ninput = nullptr;
    const void* const* in = nullptr;
    void* const* out = nullptr;
    const c8b_tag* in_tags = nullptr;
    int n_in_tags = 0;
    c8b_tag* out_tags = nullptr;
    int out_tag_cap = 0;
    uint8_t* msg = nullptr;
    int msg_cap = 0;
    // results
    int consumed = 0, produced = 0, n_out_tags = 0, msg_bytes = 0;

    const c8b_tag* tag_at(int idx) const
    {
        for (int k = 0; k < n_in_tags; k++)
            if (in_tags[k].port == 0 && in_tags[k].idx == idx) return &in_tags[k];
        return nullptr;
    }
    c8b_tag* new_tag(int idx)
    {
        if (n_out_tags >= out_tag_cap) return nullptr;
        c8b_tag* t = &out_tags[n_out_tags++];
        memset(t, 0, sizeof(*t));
        t->idx = idx;
        return t;
    }
};

struct SyncRes { int ok, mIndex; float rad, snr, rssi; };
struct SignalRes { int ok, mcs, len, nsamp; float chan[128]; };

// ---- trigger (lib/trigger_impl.cc:59-117): float preac -> flag bytes; the FSM state lives with the backend ----------
template <class Ops>
int trigger_work(Ops& ops, WorkIO& io)
{
    const int n = std::min(io.noutput, io.ninput[0]);            // forecast 1:1 (:53-57)
    if (n > 0) {
        const int rc = ops.trigger(static_cast<const float*>(io.in[0]), n, static_cast<uint8_t*>(io.out[0]));
        if (rc) return rc;
    }
    io.consumed = io.produced = n;
    return 0;
}

// ---- sync (lib/sync_impl.cc:61-153) ----------------------------------------------------------------------------------
struct SyncState {
    int sync = 0;                  // 0 SYNC_S_IDLE, 1 SYNC_S_SYNC
    float conj[2] = { 0.f, 0.f };  // d_conjMultiAvg, latched at 0x02 flags (:84-87)
};

template <class Ops>
int sync_work(Ops& ops, SyncState& s, WorkIO& io)
{
    const uint8_t* trig = static_cast<const uint8_t*>(io.in[0]);
    const float* conj = static_cast<const float*>(io.in[1]);
    const float* sig = static_cast<const float*>(io.in[2]);
    uint8_t* out = static_cast<uint8_t*>(io.out[0]);
    const int nProc = std::min(std::min(io.noutput, io.ninput[0]), std::min(io.ninput[1], io.ninput[2]));
    if (s.sync == 0) {                                            // :73-92
        int i;
        for (i = 0; i < nProc; i++) {
            out[i] = 0;
            if (trig[i] & 0x01) { s.sync = 1; break; }
            else if (trig[i] & 0x02) { s.conj[0] = conj[2 * i]; s.conj[1] = conj[2 * i + 1]; }
        }
        io.consumed = io.produced = i;
        return 0;
    }
    if (nProc < 240) return 0;                                    // :94,148-152: wait for SYNC_MAX_BUF_LEN samples
    SyncRes r;
    const int rc = ops.sync_at(sig, s.conj, &r);                  // ltf_autoCorrelation + ltf_cfo on the device
    if (rc) return rc;
    memset(out, 0, 111);                                          // :98
    if (r.ok) {
        out[r.mIndex] = 0x01;                                     // :123
        c8b_tag* t = io.new_tag(r.mIndex);                        // :124-136
        if (!t) return C8B_ERR_FULL;
        t->f.rad = r.rad; t->f.snr = r.snr; t->f.rssi = r.rssi;
    }
    s.sync = 0;
    io.consumed = io.produced = 111;                              // SYNC_MAX_RES_LEN (:145-146)
    return 0;
}

// ---- signal / signal2 (lib/signal_impl.cc:62-206, lib/signal2_impl.cc:63-212) ----------------------------------------
struct SignalState {
    int st = 0;                    // 0 S_TRIGGER, 1 S_DEMOD, 2 S_COPY, 3 S_PAD
    float rad = 0.f, snr = 0.f, rssi = 0.f;
    int seq = 0, nSample = 0, nCopied = 0;
};

template <class Ops>
int signal_work(Ops& ops, SignalState& s, int nant, WorkIO& io)
{
    const uint8_t* sync = static_cast<const uint8_t*>(io.in[0]);
    const float* in1 = static_cast<const float*>(io.in[1]);
    const float* in2 = nant == 2 ? static_cast<const float*>(io.in[2]) : nullptr;
    float* out1 = static_cast<float*>(io.out[0]);
    float* out2 = nant == 2 ? static_cast<float*>(io.out[1]) : nullptr;
    int nProc = std::min(io.ninput[0], io.ninput[1]);
    if (nant == 2) nProc = std::min(nProc, io.ninput[2]);
    int nUsed = 0, nPassed = 0;
    const float* rot0 = nullptr;                                  // S_COPY samples the backend already corrected together with the L-SIG
    const float* rot1 = nullptr;
    int nrot = 0;

    if (s.st == 0) {                                              // :75-106
        int i;
        for (i = 0; i < nProc; i++) {
            if (sync[i]) {
                const c8b_tag* t = io.tag_at(i);
                if (t) { s.rad = t->f.rad; s.snr = t->f.snr; s.rssi = t->f.rssi; s.st = 1; }
                else { printf("ieee80211 signal%s, error: input sync with no tag.\n", nant == 2 ? "2" : ""); i++; }
                break;
            }
        }
        nUsed += i;
    }
    if (s.st == 1) {                                              // :108-162
        if (nProc - nUsed >= 224) {
            SignalRes r;
            // both blocks read the L-SIG on antenna 0; the samples S_COPY would hand on in this same call ride along
            nrot = std::max(0, std::min(io.noutput, nProc - nUsed - 224));
            const int rc = ops.signal_at(in1 + 2 * (size_t)nUsed, in2 ? in2 + 2 * (size_t)nUsed : nullptr, nrot, s.rad, &r, &rot0, &rot1);
            if (rc) return rc;
            if (r.ok) {
                s.nSample = r.nsamp; s.nCopied = 0;
                s.seq++;
                if (s.seq >= 1000000000) s.seq = 0;
                c8b_tag* t = io.new_tag(0);                       // nitems_written(0): nothing produced yet in this call
                if (!t) return C8B_ERR_FULL;
                t->seq = s.seq;
                t->f.rad = s.rad;
                t->f.cfo_hz = s.rad * 3183098.8618379068f;        // :137
                t->f.snr = s.snr; t->f.rssi = s.rssi;
                t->f.l_mcs = r.mcs; t->f.l_len = r.len; t->f.nsamp = r.nsamp;
                t->nvec = 64;
                memcpy(t->vec, r.chan, sizeof(float) * 128);
                s.st = 2;
                nUsed += 224;
            } else {
                s.st = 0;
                nUsed += 80;
            }
        }
    }
    if (s.st == 2) {                                              // :164-192
        int nGen = std::min(io.noutput, nProc - nUsed);
        const int left = s.nSample - s.nCopied;
        const bool last = !(nGen < left);
        if (last) nGen = left;
        if (nGen > 0) {
            if (rot0 && s.nCopied == 0 && nGen <= nrot && (!in2 || rot1)) {
                memcpy(out1, rot0, sizeof(float) * 2 * (size_t)nGen);
                if (in2) memcpy(out2, rot1, sizeof(float) * 2 * (size_t)nGen);
            } else {
                const int rc = ops.cfo_copy(in1 + 2 * (size_t)nUsed, in2 ? in2 + 2 * (size_t)nUsed : nullptr, out1, out2, nGen, s.nCopied, s.rad);
                if (rc) return rc;
            }
        }
        s.nCopied += nGen;
        nUsed += nGen;
        nPassed += nGen;
        if (last) s.st = 3;
    }
    if (s.st == 3) {                                              // :194-201
        if (io.noutput - nPassed >= 320) {
            memset(out1 + 2 * (size_t)nPassed, 0, sizeof(float) * 2 * 320);
            if (out2) memset(out2 + 2 * (size_t)nPassed, 0, sizeof(float) * 2 * 320);
            s.st = 0;
            nPassed += 320;
        }
    }
    io.consumed = nUsed;
    io.produced = nPassed;
    return 0;
}

// ---- demod / demod2 (lib/demod_impl.cc:59-342, lib/demod2_impl.cc:58-348) ---------------------------------------------
// Two halves that share a call: the INPUT side reads the tag (RDTAG), gathers the frame's nsamp + 320 samples and submits
// them to the backend (header states + per-symbol demod on the device, asynchronous); the OUTPUT side collects finished
// frames in order and emits tag + soft bits as the output space allows.  Up to two frames are in flight, so the device
// works on frame k while the samples of frame k + 1 are gathered.
struct DemodState {
    int st = 0;                    // input side: 0 RDTAG, 1 gather (FORMAT..CLEAN of the reference)
    c8b_tag tag;                   // signal's tag of the frame being gathered
    int need = 0, have = 0;        // samples of the frame: nsamp + 320
    std::vector<float> buf[2];     // gathered samples per antenna, behind 224 zeros (the L-LTF/L-SIG part signal consumed)
    c8b_tag qtag[2];               // signal's tags of the frames in flight, oldest first
    int nq = 0;
    bool emitting = false;         // output side: a collected frame is being handed on
    std::vector<float> llr;
    c8b_frame f;
    c8b_tag etag;
    int emitted = 0, total = 0;
    bool tagPending = false;
};

template <class Ops>
int demod_work(Ops& ops, DemodState& s, int nant, WorkIO& io)
{
    int nProc = io.ninput[0];
    if (nant == 2) nProc = std::min(nProc, io.ninput[1]);
    const bool canRead = s.nq < 2 && nProc > 0 && (s.st == 1 || io.tag_at(0) != nullptr);
    // ---- output side: the oldest frame in flight, when it is done (waited for only if this call has nothing else to do) ----
    while (!s.emitting && s.nq > 0) {
        const float* soft = nullptr;
        int nsoft = 0;
        const int rc = ops.demod_collect(!canRead, &s.f, &soft, &nsoft);
        if (rc < 0) return rc;
        if (rc == 0) break;
        s.etag = s.qtag[0];
        s.qtag[0] = s.qtag[1];
        s.nq--;
        if (s.f.status == C8B_ST_OK) s.total = s.f.total;
        else if (s.f.status == C8B_ST_NDP) s.total = 1024;        // :251-257
        else continue;                                            // DEMOD_S_CLEAN: dropped, nothing leaves the block
        s.llr.assign((size_t)std::max(s.total, 1024), 0.f);
        memcpy(s.llr.data(), soft, sizeof(float) * (size_t)std::min(std::max(s.total, s.f.status == C8B_ST_NDP ? 256 : 0), nsoft));
        s.emitted = 0; s.tagPending = true; s.emitting = true;
    }
    if (s.emitting && io.noutput > 0) {                           // tag at the first soft bit (:224-263), then `total` floats
        if (s.tagPending) {
            c8b_tag* t = io.new_tag(0);
            if (!t) return C8B_ERR_FULL;
            t->f = s.f;
            t->f.snr = s.etag.f.snr; t->f.rssi = s.etag.f.rssi; t->f.cfo_hz = s.etag.f.cfo_hz;
            t->seq = s.etag.seq;
            if (s.f.status == C8B_ST_NDP) {
                t->f.total = 1024; t->f.trellis = 0;
                t->nvec = 128;
                memcpy(t->vec, s.llr.data(), sizeof(float) * 256);    // tag "mu2x1chan" (:238-249)
                std::fill(s.llr.begin(), s.llr.end(), 0.f);
            }
            s.tagPending = false;
        }
        const int n = std::min(io.noutput, s.total - s.emitted);
        memcpy(io.out[0], s.llr.data() + s.emitted, sizeof(float) * (size_t)n);
        s.emitted += n;
        io.produced = n;
        if (s.emitted >= s.total) s.emitting = false;
    }
    // ---- input side ----
    if (s.nq >= 2 || nProc <= 0) return 0;
    if (s.st == 0) {                                              // DEMOD_S_RDTAG (:72-103): a frame starts at a tagged item
        const c8b_tag* t = io.tag_at(0);
        if (!t) return 0;
        s.tag = *t;
        s.need = t->f.nsamp + 320;                                // :93
        s.have = 0;
        for (int a = 0; a < nant; a++) s.buf[a].assign(2 * (size_t)(224 + s.need), 0.f);
        s.st = 1;
    }
    const int n = std::min(nProc, s.need - s.have);
    for (int a = 0; a < nant; a++)
        memcpy(s.buf[a].data() + 2 * (size_t)(224 + s.have), io.in[a], sizeof(float) * 2 * (size_t)n);
    s.have += n;
    io.consumed = n;
    if (s.have < s.need) return 0;
    // the whole frame is here: header states + per-symbol demod in one go on the device
    c8b_frame f;
    memset(&f, 0, sizeof(f));
    f.status = C8B_ST_OK;
    f.sync_idx = 0; f.rad = 0.f;                                  // the input is signal's CFO-corrected copy
    f.snr = s.tag.f.snr; f.rssi = s.tag.f.rssi; f.cfo_hz = s.tag.f.cfo_hz;
    f.l_mcs = s.tag.f.l_mcs; f.l_len = s.tag.f.l_len; f.nsamp = s.tag.f.nsamp;
    const int rc = ops.demod_submit(nant, s.buf[0].data(), nant == 2 ? s.buf[1].data() : nullptr, 224 + s.need, &f, s.tag.vec);
    if (rc) return rc;
    s.qtag[s.nq++] = s.tag;
    s.st = 0;
    return 0;
}

// ---- decode (lib/decode_impl.cc:60-162) --------------------------------------------------------------------------------
// The block has no stream output, only messages: a frame whose soft bits are all here is SUBMITTED to the backend and the
// call returns; its MPDUs are published by a later call (in submission order), or by a call with no input -- what the shell's
// stop() / a drained scheduler issues.  The GPU round trip of one frame thereby overlaps the gathering of the next.
struct DecodeState {
    int st = 0;                    // 0 IDLE, 1 gather (DECODE), 2 CLEAN
    c8b_frame f;
    int total = 0, have = 0;
    std::vector<float> llr;
    int inflight = 0;              // frames submitted and not yet published
};

template <class Ops>
int decode_publish(Ops& ops, DecodeState& s, WorkIO& io, bool wait)
{
    while (s.inflight > 0) {
        c8b_frame f;
        const uint8_t* pdu = nullptr;
        // room for the record first: a finished frame must not be taken off the queue and then dropped
        if (io.msg_cap - io.msg_bytes < 2 * 4400 || io.n_out_tags >= io.out_tag_cap) return io.msg_bytes > 0 || io.n_out_tags > 0 ? 0 : C8B_ERR_FULL;
        const int rc = ops.decode_collect(wait, &f, &pdu);
        if (rc < 0) return rc;
        if (rc == 0) break;
        s.inflight--;
        if (c8b_tag* t = io.new_tag(-1)) { t->port = -1; t->f = f; }      // report of the finished frame (debug lines of the shell)
        if (f.npdu > 0 && f.pdu_bytes > 0) {                     // [fmt][len lo][len hi][MPDU][mcs] per CRC-passing MPDU (:512-516)
            memcpy(io.msg + io.msg_bytes, pdu, (size_t)f.pdu_bytes);
            io.msg_bytes += f.pdu_bytes;
        }
    }
    return 0;
}

template <class Ops>
int decode_work(Ops& ops, DecodeState& s, WorkIO& io)
{
    const int nProc = io.ninput[0];
    {
        const int rc = decode_publish(ops, s, io, nProc <= 0);    // nothing to read: wait for what is in flight
        if (rc) return rc;
    }
    if (s.st == 0) {                                              // :67-127
        const c8b_tag* t = io.tag_at(0);
        if (!t || nProc <= 0) return 0;
        s.f = t->f;
        s.total = t->f.total; s.have = 0;
        if (s.f.len > 4095 || s.f.trellis > 32782) { s.st = 2; }  // :93-97
        else if (s.f.trellis == 0) {                              // VHT NDP channel report (:100-121)
            s.st = 2;
            const int rc = decode_publish(ops, s, io, true);      // messages stay in stream order
            if (rc) return rc;
            const int n = 3 + 1024;
            if (io.msg_bytes + n > io.msg_cap) return C8B_ERR_FULL;
            uint8_t* m = io.msg + io.msg_bytes;
            m[0] = 20; m[1] = 1024 % 256; m[2] = 1024 / 256;      // C8P_F_VHT_CHAN, sizeof(float)*256
            memcpy(m + 3, t->vec, 1024);
            io.msg_bytes += n;
        } else {
            s.llr.assign((size_t)std::max(s.total, 0), 0.f);
            s.st = 1;
        }
        return 0;
    }
    const int n = std::min(nProc, s.total - s.have);
    if (s.st == 1) {
        memcpy(s.llr.data() + s.have, io.in[0], sizeof(float) * (size_t)n);
        s.have += n;
        io.consumed = n;
        if (s.have < s.total) return 0;
        s.f.status = C8B_ST_OK; s.f.llr_off = 0; s.f.npdu = 0; s.f.pdu_bytes = 0;
        if (s.inflight >= 3) {                                    // every staging slot but one is busy: wait for the oldest frames first
            const int rp = decode_publish(ops, s, io, true);
            if (rp) return rp;
        }
        if (s.inflight >= 4) { s.have = s.total; return 0; }      // (no room to publish into: the caller comes back with this frame still gathered)
        const int rc = ops.decode_submit(&s.f, s.llr.data(), s.total);
        if (rc) return rc;
        s.inflight++;
        s.st = 0;
        return 0;
    }
    s.have += n;                                                  // DECODE_S_CLEAN (:141-157)
    io.consumed = n;
    if (s.have >= s.total) s.st = 0;
    return 0;
}

}  // namespace c8b_blocks
