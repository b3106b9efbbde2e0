// TEST INFRASTRUCTURE (oracle/_ref) -- NOT PART OF THE PRODUCT.
//
// Runs the reference's seven receive blocks themselves: /root/reference/lib/{trigger,sync,signal,signal2,demod,demod2,
// decode}_impl.cc and cloud80211phy.cc are compiled UNMODIFIED, from where they lie, against the miniature GNU Radio
// runtime of tests/gr_mock/include (gr::block, io_signature, pmt, gr::fft, boost::crc_32_type: the image has no GNU
// Radio / FFTW / Boost) and linked with this file into oracle/_ref/libgr80211_ref.so (oracle/Makefile, target `ref`).
// Nothing here restates a block: this file is the SCHEDULER -- it wires trigger -> sync -> signal[2] -> demod[2] ->
// decode like examples/rx.grc / rx2.grc, calls forecast() / general_work() the way GNU Radio's block executor does
// (gnuradio-runtime block_executor.cc: output space first, forecast halving, every available input item offered), moves
// the items and tags along the edges, and hands the streams, tags and messages to the caller through extern "C" doors.
// presiso (stock GNU Radio blocks in the reference's flowgraph, examples/presiso.grc) comes from oracle_rx's
// orx_presiso.
//
// Used (a) as the chain-level oracle the restatement oracle_rx.cc and the CUDA path are pinned to
// (tests/test_oracle_vs_ref.py, tests/test_ref_chain.py), (b) as bench.py's `--impl reference` arm and cpu_baseline
// ("kind": "reference").  Never linked into the product library.
#include <gnuradio/ieee80211/decode.h>
#include <gnuradio/ieee80211/demod.h>
#include <gnuradio/ieee80211/demod2.h>
#include <gnuradio/ieee80211/signal.h>
#include <gnuradio/ieee80211/signal2.h>
#include <gnuradio/ieee80211/sync.h>
#include <gnuradio/ieee80211/trigger.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

extern "C" void orx_presiso(const float* iq, int n, float* preac, float* preconj);   // oracle/oracle_rx.cc

using gr::mock::edge;

namespace {

struct Rng {
    uint32_t s = 1;
    uint32_t next() { s = s * 1664525u + 1013904223u; return s >> 8; }
};

// GNU Radio's default buffer is 64 KiB per edge (gnuradio-runtime flat_flowgraph.cc: GR_FIXED_BUFFER_SIZE)
int default_buf_items(int item) { return 65536 / item; }

struct Chain {
    int nant = 1;
    std::shared_ptr<gr::block> blk[5];                 // trigger, sync, signal[2], demod[2], decode
    std::vector<std::unique_ptr<edge>> edges;
    edge *srcAc = nullptr, *srcConj = nullptr, *srcSigSync = nullptr, *srcSig[2] = { nullptr, nullptr };
    bool record = false, randomCalls = false;
    int maxCall = 0;                                   // 0: GNU Radio's buffer sizes
    Rng rng;
    // recorded output streams (record == true): trigger u8, sync u8, signal c64 x nant, demod f32
    std::vector<char> rec[5];
    uint64_t calls = 0;
    std::vector<char> obuf[2];

    edge* make_edge(int item) { edges.emplace_back(new edge); edges.back()->item = item; return edges.back().get(); }
    void connect(gr::block& a, int pa, gr::block& b, int pb)
    {
        edge* e = make_edge(a.output_signature()->sizeof_stream_item(pa));
        a.mock_out.at(pa).push_back(e);
        b.mock_in.at(pb) = e;
    }
    edge* source(int item, gr::block& b, int pb)
    {
        edge* e = make_edge(item);
        b.mock_in.at(pb) = e;
        return e;
    }

    Chain(int nant_, int mupos, int mugid, bool dbg) : nant(nant_)
    {
        using namespace gr::ieee80211;
        blk[0] = trigger::make();
        blk[1] = sync::make();
        if (nant == 2) { blk[2] = signal2::make(); blk[3] = demod2::make(); }
        else { blk[2] = signal::make(); blk[3] = demod::make(mupos, mugid); }
        blk[4] = decode::make(dbg);
        srcAc = source(4, *blk[0], 0);
        connect(*blk[0], 0, *blk[1], 0);
        srcConj = source(8, *blk[1], 1);
        srcSigSync = source(8, *blk[1], 2);
        connect(*blk[1], 0, *blk[2], 0);
        srcSig[0] = source(8, *blk[2], 1);
        if (nant == 2) srcSig[1] = source(8, *blk[2], 2);
        for (int a = 0; a < nant; a++) connect(*blk[2], a, *blk[3], a);
        connect(*blk[3], 0, *blk[4], 0);
    }

    static void append(edge* e, const void* p, size_t items)
    {
        const char* c = (const char*)p;
        e->data.insert(e->data.end(), c, c + items * (size_t)e->item);
        e->nwritten += items;
    }
    void feed(const float* preac, const float* preconj, const float* s0, const float* s1, size_t n)
    {
        append(srcAc, preac, n);
        append(srcConj, preconj, n);
        append(srcSigSync, s0, n);
        append(srcSig[0], s0, n);
        if (nant == 2) append(srcSig[1], s1, n);
    }

    static void consume(edge* e, size_t items)
    {
        e->head += items * (size_t)e->item;
        e->nread += items;
        if (e->head > (1u << 20) && e->head * 2 > e->data.size()) {
            e->data.erase(e->data.begin(), e->data.begin() + (ptrdiff_t)e->head);
            e->head = 0;
        }
        if (e->tags.size() > 48) {
            size_t k = 0;
            for (auto& t : e->tags) if (t.offset >= e->nread) e->tags[k++] = t;
            e->tags.resize(k);
        }
    }

    // One pass of the block executor over block b.  `flush` lifts the forecast (end of a finite capture: GNU Radio would
    // wait for more items; the harness offers what is left so the last frame completes).  Returns whether anything moved.
    bool call(int bi, bool flush, bool big)
    {
        gr::block& b = *blk[bi];
        const int nin = (int)b.mock_in.size(), nout = (int)b.mock_out.size();
        size_t availMin = (size_t)-1, availMax = 0;
        std::vector<size_t> avail(nin);
        for (int k = 0; k < nin; k++) {
            avail[k] = b.mock_in[k]->avail();
            const size_t cap = (size_t)(maxCall > 0 ? maxCall : default_buf_items(b.mock_in[k]->item));
            avail[k] = std::min(avail[k], cap);
            availMin = std::min(availMin, avail[k]);
            availMax = std::max(availMax, avail[k]);
        }
        int noutput;
        if (nout > 0) {
            // space left in the (bounded) output buffer: capacity - 1 - what the reader has not consumed yet
            long space = 1L << 30;
            for (int k = 0; k < nout; k++)
                for (edge* e : b.mock_out[k]) {
                    const long cap = maxCall > 0 ? maxCall : default_buf_items(e->item);
                    space = std::min(space, cap - 1 - (long)e->avail());
                }
            if (space <= 0) return false;                                     // BLKD_OUT
            noutput = (int)space;
        } else {
            if (availMax == 0) return false;
            noutput = (int)availMax;                                          // sink: relative rate 1
        }
        const bool rnd = randomCalls && !big;
        if (rnd) noutput = std::min<long>(noutput, 1 + (long)(rng.next() % (uint32_t)std::max(maxCall, 1)));
        gr_vector_int req(nin), ninput(nin);
        // forecast halving (block_executor.cc try_again).  End of a finite capture (`flush`): signal / demod / decode size their
        // work from ninput_items, so they may be offered what is left without the forecast (demod needs noutput > nCBPS,
        // which the halving against a short input tail never reaches; decode's forecast asks for 160 items more than it
        // gets); trigger and sync take noutput_items as their input count and keep the forecast.
        const bool lift = flush && bi >= 2 && availMin > 0;
        for (;;) {
            b.forecast(noutput, req);
            bool ok = lift;
            if (!ok) { ok = true; for (int k = 0; k < nin; k++) ok &= (size_t)req[k] <= avail[k]; }
            if (ok) break;
            if (noutput > 1) { noutput /= 2; continue; }
            return false;                                                     // BLKD_IN
        }
        const int extra = rnd ? (int)(rng.next() % 64u) : 0;
        for (int k = 0; k < nin; k++)
            ninput[k] = rnd ? (int)std::min<size_t>(avail[k], (size_t)std::max(req[k], noutput) + extra) : (int)avail[k];
        gr_vector_const_void_star in(nin);
        for (int k = 0; k < nin; k++) in[k] = b.mock_in[k]->data.data() + b.mock_in[k]->head;
        gr_vector_void_star out(nout);
        for (int k = 0; k < nout; k++) {
            const size_t bytes = (size_t)noutput * b.output_signature()->sizeof_stream_item(k);
            if (obuf[k].size() < bytes) obuf[k].resize(bytes);
            memset(obuf[k].data(), 0, bytes);                                 // the pad samples signal leaves unwritten read as zeros
            out[k] = obuf[k].data();
        }
        b.mock_consumed = 0;
        const size_t tagsBefore = b.mock_tags_added.size(), msgBefore = b.mock_messages.size();
        calls++;
        const int produced = b.general_work(noutput, ninput, in, out);
        if (produced < 0 || produced > noutput || b.mock_consumed < 0) { fprintf(stderr, "ref_chain: %s: bad accounting (%d of %d produced, %d consumed)\n", b.name().c_str(), produced, noutput, b.mock_consumed); abort(); }
        for (int k = 0; k < nin; k++) {
            if ((size_t)b.mock_consumed > b.mock_in[k]->avail()) { fprintf(stderr, "ref_chain: %s consumed more than available\n", b.name().c_str()); abort(); }
            consume(b.mock_in[k], (size_t)b.mock_consumed);
        }
        for (int k = 0; k < nout; k++) {
            for (edge* e : b.mock_out[k]) append(e, obuf[k].data(), (size_t)produced);
            b.mock_written[k] += produced;
            if (record) {
                const int slot = bi < 2 ? bi : bi == 2 ? 2 + k : 4;
                rec[slot].insert(rec[slot].end(), obuf[k].begin(), obuf[k].begin() + (ptrdiff_t)((size_t)produced * b.output_signature()->sizeof_stream_item(k)));
            }
        }
        const bool moved = b.mock_consumed > 0 || produced > 0 || b.mock_tags_added.size() != tagsBefore || b.mock_messages.size() != msgBefore;
        if (!record) {
            if (b.mock_tags_added.size() > 4096) b.mock_tags_added.clear();
        }
        return moved;
    }

    // run until nothing moves for three full rounds (state changes with 0 consumed / 0 produced need up to two more calls;
    // after a quiet round the pseudo-random call sizes give way to the largest ones, so nothing is left waiting for a lucky draw)
    void run(bool flush)
    {
        for (int idle = 0; idle < 3;) {
            bool moved = false;
            for (int bi = 0; bi < 5; bi++) {
                int quiet = 0;
                for (int guard = 0; guard < 1 << 20 && quiet < 3; guard++) {
                    if (call(bi, flush && idle > 0, idle > 0)) { moved = true; quiet = 0; }
                    else quiet++;
                }
            }
            idle = moved ? 0 : idle + 1;
        }
    }
};

}  // namespace

extern "C" {

void* refchain_create(int nant, int mupos, int mugid, int ifdebug)
{
    try { return new Chain(nant, mupos, mugid, ifdebug != 0); }
    catch (const std::exception& e) { fprintf(stderr, "refchain_create: %s\n", e.what()); return nullptr; }
}
void refchain_destroy(void* h) { delete (Chain*)h; }

// seed == 0: the executor's own sizes (every available item, buffers of max_call items or GNU Radio's 64 KiB default when
// max_call == 0); seed != 0: pseudo-random call sizes up to max_call.  flush: complete the last frame of a finite capture.
int refchain_run(void* h, const float* preac, const float* preconj, const float* sig0, const float* sig1, long n,
                 unsigned seed, int max_call, int record, int flush)
{
    Chain& c = *(Chain*)h;
    c.record = record != 0;
    c.randomCalls = seed != 0;
    c.maxCall = max_call;
    c.rng.s = seed * 2654435761u + 1u;
    if (n > 0) c.feed(preac, preconj, sig0, sig1, (size_t)n);
    c.run(flush != 0);
    return 0;
}

// which: 0 trigger flags (u8), 1 sync flags (u8), 2 / 3 signal out (c64), 4 demod out (f32); returns the item count
long refchain_stream(void* h, int which, const void** p)
{
    Chain& c = *(Chain*)h;
    static const int item[5] = { 1, 1, 8, 8, 4 };
    if (which < 0 || which > 4) return -1;
    *p = c.rec[which].data();
    return (long)(c.rec[which].size() / (size_t)item[which]);
}

long refchain_ntags(void* h, int block) { return (long)((Chain*)h)->blk[block]->mock_tags_added.size(); }

// type: 0 long, 1 real, 2 c32vector (cv / ncv set, val = 0)
int refchain_tag(void* h, int block, long i, unsigned long long* offset, char* key, int keycap, int* type, double* val,
                 const float** cv, int* ncv)
{
    const gr::tag_t& t = ((Chain*)h)->blk[block]->mock_tags_added.at((size_t)i);
    *offset = t.offset;
    snprintf(key, (size_t)keycap, "%s", pmt::symbol_to_string(t.key).c_str());
    *cv = nullptr; *ncv = 0; *val = 0;
    if (auto p = dynamic_cast<const pmt::p_long*>(t.value.get())) { *type = 0; *val = (double)p->v; }
    else if (auto q = dynamic_cast<const pmt::p_real*>(t.value.get())) { *type = 1; *val = q->v; }
    else if (auto r = dynamic_cast<const pmt::p_c32v*>(t.value.get())) { *type = 2; *cv = (const float*)r->v.data(); *ncv = (int)r->v.size(); }
    else return -1;
    return 0;
}

long refchain_nmsgs(void* h) { return (long)((Chain*)h)->blk[4]->mock_messages.size(); }
long refchain_msg(void* h, long i, const unsigned char** p)
{
    const auto& m = ((Chain*)h)->blk[4]->mock_messages.at((size_t)i);
    const pmt::pmt_t blob = pmt::cdr(m.second);
    *p = (const unsigned char*)pmt::blob_data(blob);
    const long n = (long)pmt::blob_length(blob);
    if (pmt::to_long(pmt::dict_ref(pmt::car(m.second), pmt::mp("len"), pmt::from_long(-1))) != n) return -1;
    return n;
}
unsigned long long refchain_calls(void* h) { return ((Chain*)h)->calls; }
// items the blocks have not consumed yet on the five inner edges + the capture (diagnostics for the tests)
long refchain_backlog(void* h, int block, int port) { return (long)((Chain*)h)->blk[block]->mock_in.at((size_t)port)->avail(); }

// The timed CPU arm: items [0, nitems) of one capture arena are dealt to nthreads workers in contiguous runs; every worker
// owns one chain (one set of the reference's blocks) and feeds it the concatenation of its items as one stream, a
// GNU-Radio-buffer-sized piece at a time -- presiso included.  Counts the messages on decode's port.  counts[0] = messages
// (CRC-passing MPDUs + NDP reports), counts[1] = samples fed, counts[2] = general_work calls.
int refchain_bench(const float* iq, const long long* offs, const int* lens, int nitems, int nthreads, long long* counts)
{
    if (nthreads < 1) nthreads = 1;
    std::atomic<long long> msgs(0), samples(0), calls(0);
    std::atomic<int> failed(0);
    auto worker = [&](int t) {
        const int lo = (int)((long long)nitems * t / nthreads), hi = (int)((long long)nitems * (t + 1) / nthreads);
        if (lo >= hi) return;
        try {
            Chain c(1, 0, 0, false);
            c.maxCall = 0;
            const int piece = 1 << 16;
            std::vector<float> buf, ac, cj;
            // 64 samples of history keep presiso's windows whole across pieces
            std::vector<float> hist(2 * 64, 0.f);
            long long ns = 0;
            auto push = [&](const float* x, int n, bool last) {
                buf.resize((size_t)(n + 64) * 2);
                memcpy(buf.data(), hist.data(), sizeof(float) * 128);
                memcpy(buf.data() + 128, x, sizeof(float) * 2 * (size_t)n);
                ac.resize((size_t)n + 64); cj.resize(((size_t)n + 64) * 2);
                orx_presiso(buf.data(), n + 64, ac.data(), cj.data());
                if (n >= 64) memcpy(hist.data(), x + 2 * (size_t)(n - 64), sizeof(float) * 128);
                else { memmove(hist.data(), hist.data() + 2 * (size_t)n, sizeof(float) * 2 * (size_t)(64 - n)); memcpy(hist.data() + 2 * (size_t)(64 - n), x, sizeof(float) * 2 * (size_t)n); }
                c.feed(ac.data() + 64, cj.data() + 128, x, nullptr, (size_t)n);
                c.run(last);
                ns += n;
            };
            std::vector<float> stream;
            for (int i = lo; i < hi; i++) {
                const float* x = iq + 2 * offs[i];
                stream.insert(stream.end(), x, x + 2 * (size_t)lens[i]);
                while ((int)(stream.size() / 2) >= piece) {
                    push(stream.data(), piece, false);
                    stream.erase(stream.begin(), stream.begin() + 2 * (ptrdiff_t)piece);
                }
            }
            push(stream.data(), (int)(stream.size() / 2), true);
            msgs += (long long)c.blk[4]->mock_messages.size();
            samples += ns;
            calls += (long long)c.calls;
        } catch (const std::exception& e) {
            fprintf(stderr, "refchain_bench: %s\n", e.what());
            failed = 1;
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back(worker, t);
    for (auto& x : th) x.join();
    counts[0] = msgs; counts[1] = samples; counts[2] = calls;
    return failed ? -1 : 0;
}

}  // extern "C"
