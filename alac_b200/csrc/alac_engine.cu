// alac_engine.cu -- host side of libalac_b200.so: engine object, batched encode / decode
// orchestration and the extern "C" ABI declared in include/alac_b200.h.
//
// No CPU codec path exists here: every compute entry point launches the sm_100a kernels of
// alac_encode.cuh / alac_decode.cuh and fails with ALAC_B200_CUDA_ERROR if that is impossible.
#include "../../include/alac_b200.h"
#include "alac_decode.cuh"
#include "alac_encode.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <chrono>
#include <thread>
#include <vector>

using namespace alacb;

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

}  // namespace

constexpr size_t kMaxChunks = 4096;

struct alac_b200_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;     // transfer streams of the host-buffer pipeline
    cudaStream_t side = nullptr;                            // staged placement, home rank: wait-for-all + compaction
    bool finish_pending = false;                            // a deferred finish: home rank = side stream, other ranks = copy-out stream
    bool finish_far = false;
    uint64_t finish_total = 0, finish_capacity = 0, last_base = 0;
    cudaStream_t lanes[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // compute streams of the host-buffer pipeline
    cudaStream_t cur = nullptr;                             // stream the launch helpers / timers use right now
    uint64_t *h_totals = nullptr;                           // pinned: running byte / frame totals per chunk
    int32_t *h_status = nullptr;                            // pinned: per-packet status of the latest decode call
    size_t h_status_cap = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_far = nullptr, ev_far0 = nullptr;        // staged placement (trace): start of the placed call, this rank's last cross-GPU work
    std::string err;
    // encode
    DevBuf pcm, pkt_frame, pkt_samples, seg_first, seg_count, seg_stream, recs, scratch, sizes, offsets, out, state, counters, scan_tiles;
    // decode
    DevBuf d_packets, d_sizes, d_pkt_off, d_pkt_samples, d_out_frame, d_status, d_pcm, d_class, d_rank, d_perm, d_chan, d_meta, d_hdr, jobs, job_counts;
    uint32_t launches = 0;
    bool decode_configured = false;
    // several GPUs (alac_b200_engine_create_multi): subs[0] is this engine itself (the home device)
    std::vector<alac_b200_engine *> subs;
    DevBuf xchg, m_out, m_sizes, m_aux, xwords;      // home: exchange block; every sub: its block of a host-output / decode call
    uint32_t epoch = 0;
    // asynchronous forms (alac_b200_*_submit / alac_b200_wait): one call in flight on a worker thread
    std::thread worker;
    bool async_busy = false;
    int32_t async_rc = 0;
    alac_b200_enc_config async_cfg;
    std::vector<alac_b200_stream> async_streams;
    std::vector<uint8_t> async_cookie;
    // encode geometry tables of the previous call (see alac_b200_encode)
    std::vector<uint64_t> h_pkt_frame;
    std::vector<uint32_t> h_pkt_samples, h_seg_first, h_seg_count, h_seg_stream;
    std::vector<alac_b200_stream> tab_streams;
    uint32_t tab_F = 0, tab_K = 0;
    uint64_t tab_sample_sum = 0;     // sum of the streams' lengths (the exact output bound of the table)
    bool tables_valid = false;
    // per-kernel timers: (start, stop) event pairs, grown on demand, reused across calls
    std::vector<cudaEvent_t> timers;
    size_t timers_used = 0;
    std::vector<cudaEvent_t> t_mid;     // (before, between, after) of every two-kernel launch: search | final, entropy | finish
    cudaEvent_t event_on(cudaStream_t s)
    {
        if (timers_used == timers.size()) {
            cudaEvent_t ev = nullptr;
            cudaEventCreate(&ev);
            timers.push_back(ev);
        }
        cudaEvent_t ev = timers[timers_used++];
        cudaEventRecord(ev, s);
        return ev;
    }
    cudaEvent_t new_event()         // not recorded yet
    {
        if (timers_used == timers.size()) {
            cudaEvent_t ev = nullptr;
            cudaEventCreate(&ev);
            timers.push_back(ev);
        }
        return timers[timers_used++];
    }
    cudaEvent_t timer()
    {
        if (timers_used == timers.size()) {
            cudaEvent_t ev = nullptr;
            cudaEventCreate(&ev);
            timers.push_back(ev);
        }
        cudaEvent_t ev = timers[timers_used++];
        cudaEventRecord(ev, cur ? cur : stream);
        return ev;
    }
};

#define CU_CHECK(eng, call)                                                                        \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            (eng)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                      \
            return e__ == cudaErrorMemoryAllocation ? ALAC_B200_MEM_ERROR : ALAC_B200_CUDA_ERROR;  \
        }                                                                                          \
    } while (0)

// codec/ALACEncoder.cu:97-107 sChannelMaps: element tags per channel index, 3 bits each
static const uint32_t kChannelMaps[8] = {
    ID_SCE,
    ID_CPE,
    (ID_CPE << 3) | (ID_SCE),
    (ID_SCE << 9) | (ID_CPE << 3) | (ID_SCE),
    (ID_CPE << 9) | (ID_CPE << 3) | (ID_SCE),
    (ID_SCE << 15) | (ID_CPE << 9) | (ID_CPE << 3) | (ID_SCE),
    (ID_SCE << 18) | (ID_SCE << 15) | (ID_CPE << 9) | (ID_CPE << 3) | (ID_SCE),
    (ID_SCE << 21) | (ID_CPE << 15) | (ID_CPE << 9) | (ID_CPE << 3) | (ID_SCE)};

// codec/ALACAudioTypes.h:115-125
static const uint32_t kLayoutTags[8] = {(100u << 16) | 1, (101u << 16) | 2, (113u << 16) | 3, (116u << 16) | 4,
                                        (120u << 16) | 5, (124u << 16) | 6, (142u << 16) | 7, (127u << 16) | 8};

// how many chunks the host-buffer pipeline cuts a call into (ALAC_B200_PIPELINE_CHUNKS overrides)
static uint32_t pipeline_chunks()
{
    static const uint32_t n = [] {
        const char *v = getenv("ALAC_B200_PIPELINE_CHUNKS");
        const long k = v ? atol(v) : 5;
        return (uint32_t)(k < 1 ? 1 : k > 64 ? 64 : k);
    }();
    return n;
}

// Share of the packets that pipeline chunk i of n gets.  Every kernel here is latency-bound (a packet is one serial
// chain), so a chunk's kernels take about as long however small it is; what the pipeline cannot hide is the
// kernel time of the LAST chunk on encode (after the last PCM byte arrived) and of the FIRST chunk on decode
// (before the first PCM byte can leave).  Those chunks are therefore the small ones: linear weights n, n-1, .. 1.
static uint64_t tapered_chunk(uint64_t total, uint32_t n, uint32_t i, bool small_last)
{
    // ALAC_B200_PIPELINE_WEIGHTS="w0,w1,..." (encode order; decode uses it reversed) overrides the linear taper
    static const std::vector<uint64_t> custom = [] {
        std::vector<uint64_t> w;
        const char *v = getenv("ALAC_B200_PIPELINE_WEIGHTS");
        while (v && *v) {
            char *end = nullptr;
            const long long k = strtoll(v, &end, 10);
            if (end == v) break;
            w.push_back((uint64_t)std::max<long long>(1, k));
            v = (*end == ',') ? end + 1 : end;
        }
        return w;
    }();
    if (!custom.empty()) {
        const uint32_t m = (uint32_t)custom.size();
        uint64_t wsum = 0;
        for (uint64_t w : custom) wsum += w;
        const uint32_t k = std::min(i, m - 1);
        return std::max<uint64_t>(1, (total * custom[small_last ? k : m - 1 - k] + wsum - 1) / wsum);
    }
    const uint64_t wsum = (uint64_t)n * (n + 1) / 2;
    const uint64_t w = small_last ? (n - i) : (i + 1);
    return std::max<uint64_t>(1, (total * w + wsum - 1) / wsum);
}

// packets per search launch: bounds the Golomb-slab scratch (ALAC_B200_CHUNK_PACKETS overrides)
static uint64_t scratch_chunk_packets()
{
    static const uint64_t n = [] {
        const char *v = getenv("ALAC_B200_CHUNK_PACKETS");
        const long long k = v ? atoll(v) : 0;
        return (uint64_t)(k <= 0 ? 0 : k < 1024 ? 1024 : k);
    }();
    return n;
}

static bool valid_depth(uint32_t d) { return d == 16 || d == 20 || d == 24 || d == 32; }
static uint32_t bytes_per_sample(uint32_t d) { return d == 16 ? 2u : d == 32 ? 4u : 3u; }

static bool valid_cfg(const alac_b200_enc_config *c)
{
    return c && c->channels >= 1 && c->channels <= 8 && valid_depth(c->bit_depth) && c->frame_size >= 1 &&
           c->frame_size <= 16384;
}

// element walk of ALACEncoder::Encode (codec/ALACEncoder.cu:973-1057; >2 channels per sChannelMaps)
static void build_layout(const alac_b200_enc_config *c, EncLayout &L, uint32_t &mono_mask, uint32_t &pair_mask)
{
    memset(&L, 0, sizeof(L));
    L.channels = c->channels;
    L.frame_size = c->frame_size;
    L.fast_mode = c->fast_mode ? 1u : 0u;
    mono_mask = pair_mask = 0;
    uint32_t ch = 0, e = 0, chain = 0, mono_tag = 0, pair_tag = 0, lfe_tag = 0;
    while (ch < c->channels) {
        const uint32_t tag = (kChannelMaps[c->channels - 1] >> (ch * 3)) & 7u;
        L.elem_tag[e] = (uint8_t)tag;
        L.elem_chan[e] = (uint8_t)ch;
        L.elem_chain[e] = (uint8_t)chain;
        if (tag == ID_CPE) {
            L.elem_inst[e] = (uint8_t)pair_tag++;
            pair_mask |= 1u << e;
            ch += 2;
            chain += 2;
        } else {
            L.elem_inst[e] = (uint8_t)(tag == ID_LFE ? lfe_tag++ : mono_tag++);
            mono_mask |= 1u << e;
            ch += 1;
            chain += 1;
        }
        e++;
    }
    L.elems_per_packet = e;
    L.chains_per_packet = chain;
}

static inline void put_be32(uint8_t *p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }

extern "C" {

const char *alac_b200_version(void) { return "alac_b200 0.1 (sm_100a)"; }

int32_t alac_b200_engine_create(int32_t device, alac_b200_engine **out_engine)
{
    if (!out_engine) return ALAC_B200_PARAM_ERROR;
    *out_engine = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return ALAC_B200_CUDA_ERROR;
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) return ALAC_B200_CUDA_ERROR;
    }
    if (device >= count) return ALAC_B200_PARAM_ERROR;
    if (cudaSetDevice(device) != cudaSuccess) return ALAC_B200_CUDA_ERROR;
    alac_b200_engine *e = new alac_b200_engine();
    e->device = device;
    if (cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete e;
        return ALAC_B200_CUDA_ERROR;
    }
    e->stream = e->own_stream;
    bool lanes_ok = true;
    for (auto &ln : e->lanes) lanes_ok = lanes_ok && cudaStreamCreateWithFlags(&ln, cudaStreamNonBlocking) == cudaSuccess;
    if (!lanes_ok || cudaStreamCreateWithFlags(&e->copy_in, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&e->copy_out, cudaStreamNonBlocking) != cudaSuccess ||
        [&] {   // the side stream closes the gaps of a staged job while this GPU already decodes: highest priority, so its CTAs
                // are placed as soon as slots free up instead of after the decode kernels' queued CTAs
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);
            return cudaStreamCreateWithPriority(&e->side, cudaStreamNonBlocking, hi);
        }() != cudaSuccess ||
        cudaHostAlloc(&e->h_totals, kMaxChunks * sizeof(uint64_t), cudaHostAllocDefault) != cudaSuccess) {
        delete e;
        return ALAC_B200_CUDA_ERROR;
    }
    for (auto &ev : e->ev) {
        if (cudaEventCreate(&ev) != cudaSuccess) {
            delete e;
            return ALAC_B200_CUDA_ERROR;
        }
    }
    if (cudaEventCreate(&e->ev_far) != cudaSuccess || cudaEventCreate(&e->ev_far0) != cudaSuccess) { delete e; return ALAC_B200_CUDA_ERROR; }
    *out_engine = e;
    return ALAC_B200_OK;
}

void alac_b200_engine_destroy(alac_b200_engine *e)
{
    if (!e) return;
    if (e->worker.joinable()) e->worker.join();
    for (size_t i = 1; i < e->subs.size(); i++) alac_b200_engine_destroy(e->subs[i]);
    e->subs.clear();
    cudaSetDevice(e->device);
    // nothing may still be in flight on the engine's buffers (a deferred placed call keeps copying / compacting)
    cudaStreamSynchronize(e->stream);
    if (e->copy_in) cudaStreamSynchronize(e->copy_in);
    if (e->copy_out) cudaStreamSynchronize(e->copy_out);
    if (e->side) cudaStreamSynchronize(e->side);
    for (auto &ln : e->lanes) if (ln) cudaStreamSynchronize(ln);
    e->xchg.release(); e->m_out.release(); e->m_sizes.release(); e->m_aux.release(); e->xwords.release();
    DevBuf *bufs[] = {&e->pcm, &e->pkt_frame, &e->pkt_samples, &e->seg_first, &e->seg_count, &e->seg_stream, &e->recs,
                      &e->scratch, &e->sizes, &e->offsets, &e->out, &e->state, &e->counters, &e->scan_tiles, &e->d_packets, &e->d_sizes,
                      &e->d_pkt_off, &e->d_pkt_samples, &e->d_out_frame, &e->d_status, &e->d_pcm, &e->d_class, &e->d_rank, &e->d_perm, &e->d_chan, &e->d_meta, &e->d_hdr, &e->jobs, &e->job_counts};
    for (DevBuf *b : bufs) b->release();
    for (auto &ev : e->ev)
        if (ev) cudaEventDestroy(ev);
    if (e->ev_far) cudaEventDestroy(e->ev_far);
    if (e->ev_far0) cudaEventDestroy(e->ev_far0);
    for (auto &ev : e->timers) cudaEventDestroy(ev);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    if (e->copy_in) cudaStreamDestroy(e->copy_in);
    for (auto &ln : e->lanes) if (ln) cudaStreamDestroy(ln);
    if (e->copy_out) cudaStreamDestroy(e->copy_out);
    if (e->side) cudaStreamDestroy(e->side);
    if (e->h_totals) cudaFreeHost(e->h_totals);
    if (e->h_status) cudaFreeHost(e->h_status);
    delete e;
}

const char *alac_b200_last_error(const alac_b200_engine *e) { return e ? e->err.c_str() : "null engine"; }

int32_t alac_b200_engine_set_stream(alac_b200_engine *e, void *cuda_stream)
{
    if (!e) return ALAC_B200_PARAM_ERROR;
    const cudaStream_t ns = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : e->own_stream;
    if (ns != e->stream) e->tables_valid = false;      // cached tables were uploaded in the old stream's order
    e->stream = ns;
    return ALAC_B200_OK;
}

uint32_t alac_b200_magic_cookie(const alac_b200_enc_config *c, uint32_t max_frame_bytes, uint32_t avg_bit_rate,
                                void *out_cookie, uint32_t cap)
{
    // codec/ALACEncoder.cu:1082-1140; layout codec/ALACAudioTypes.h:162-176 (all big-endian)
    if (!valid_cfg(c) || !out_cookie) return 0;
    const uint32_t size = 24 + (c->channels > 2 ? 24u : 0u);
    if (cap < size) return 0;       // "no incomplete cookies", :1136-1139
    uint8_t *o = static_cast<uint8_t *>(out_cookie);
    memset(o, 0, size);
    put_be32(o, c->frame_size);
    o[4] = 0;
    o[5] = (uint8_t)c->bit_depth;
    o[6] = (uint8_t)kPb0;
    o[7] = (uint8_t)kMb0;
    o[8] = (uint8_t)kKb0;
    o[9] = (uint8_t)c->channels;
    o[10] = 0;
    o[11] = 255;                    // maxRun
    put_be32(o + 12, max_frame_bytes);
    put_be32(o + 16, avg_bit_rate);
    put_be32(o + 20, c->sample_rate);
    if (c->channels > 2) {
        static const uint8_t atom[12] = {0, 0, 0, 24, 'c', 'h', 'a', 'n', 0, 0, 0, 0};
        memcpy(o + 24, atom, 12);
        const uint32_t tag = kLayoutTags[c->channels - 1];      // stored native-endian, :1120
        memcpy(o + 36, &tag, 4);
    }
    return size;
}

int32_t alac_b200_parse_cookie(const void *cookie, uint32_t cookie_size, uint32_t f[11])
{
    // codec/ALACDecoder.cu:109-190
    if (!cookie || !f) return ALAC_B200_PARAM_ERROR;
    const uint8_t *p = static_cast<const uint8_t *>(cookie);
    uint32_t left = cookie_size;
    if (left >= 12 && p[4] == 'f' && p[5] == 'r' && p[6] == 'm' && p[7] == 'a') { p += 12; left -= 12; }
    if (left >= 12 && p[4] == 'a' && p[5] == 'l' && p[6] == 'a' && p[7] == 'c') { p += 12; left -= 12; }
    if (left < 24) return ALAC_B200_PARAM_ERROR;
    auto be32 = [](const uint8_t *q) { return ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3]; };
    f[0] = be32(p);
    f[1] = p[4]; f[2] = p[5]; f[3] = p[6]; f[4] = p[7]; f[5] = p[8]; f[6] = p[9];
    f[7] = ((uint32_t)p[10] << 8) | p[11];
    f[8] = be32(p + 12); f[9] = be32(p + 16); f[10] = be32(p + 20);
    if (f[1] > 0) return ALAC_B200_PARAM_ERROR;     // compatibleVersion <= kALACVersion, :153
    return ALAC_B200_OK;
}

uint64_t alac_b200_encode_bound(const alac_b200_enc_config *c, uint64_t num_sample_frames, uint64_t num_streams)
{
    if (!valid_cfg(c)) return 0;
    if (num_streams == 0) num_streams = 1;
    const uint64_t bpf = (uint64_t)bytes_per_sample(c->bit_depth) * c->channels;
    const uint64_t packets = num_sample_frames / c->frame_size + num_streams;
    // an element never exceeds its escape size: 7 + 16 + 32 header bits + raw samples; + ID_END + padding
    return num_sample_frames * bpf + packets * (7ull * c->channels + 1ull);
}

// used by ALACDecoder::fillWriteBuffer (fork signature): host -> caller's device buffer
int32_t alac_b200_copy_to_device(void *dst, const void *src, uint64_t bytes)
{
    return cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyHostToDevice) == cudaSuccess ? ALAC_B200_OK : ALAC_B200_CUDA_ERROR;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// encode
// ------------------------------------------------------------------------------------------------
// Synchronises every stream a call may have used when the call returns early (CU_CHECK, capacity errors): queued
// kernels and copies must not outlive the call -- they use engine scratch the next call reuses and write to locals
// (pinned counters, status vectors) that are destroyed on return.  Declare it AFTER those locals.
struct DrainGuard {
    alac_b200_engine *e;
    bool armed = true;
    explicit DrainGuard(alac_b200_engine *eng) : e(eng) {}
    ~DrainGuard()
    {
        if (!armed) return;
        cudaStreamSynchronize(e->stream);
        cudaStreamSynchronize(e->copy_in);
        cudaStreamSynchronize(e->copy_out);
        cudaStreamSynchronize(e->side);
        for (auto &ln : e->lanes) cudaStreamSynchronize(ln);
        e->cur = nullptr;
        e->finish_pending = false;
    }
};

static void launch_scan(alac_b200_engine *e, cudaStream_t s, const uint32_t *in, uint64_t *out, uint64_t n, uint64_t *tiles,
                        uint32_t *max_out, int chain_base, uint64_t *host_total)
{
    const uint32_t blocks = (uint32_t)std::max<uint64_t>(1, (n + kScanTile - 1) / kScanTile);
    if (blocks > 1) scan_tile_sums_kernel<<<blocks, 1024, 0, s>>>(in, n, tiles);
    scan_u32_to_u64_kernel<<<blocks, 1024, 0, s>>>(in, out, n, tiles, max_out, chain_base, host_total);
    e->launches += blocks > 1 ? 2 : 1;
}
static size_t scan_tiles_for(uint64_t n) { return (size_t)((n + kScanTile - 1) / kScanTile + 1); }

static int32_t encode_core(alac_b200_engine *e, const alac_b200_enc_config *cfg, const void *pcm,
                           uint64_t num_sample_frames, int32_t pcm_mem, const alac_b200_stream *streams,
                           uint64_t n_streams, void *packets_out, uint64_t packets_cap, uint32_t *packet_sizes,
                           uint64_t sizes_cap, int32_t out_mem, int16_t *coef_state, uint64_t *out_num_packets,
                           uint64_t *out_bytes, alac_b200_stats *stats, const alac_b200_placement *pl, uint64_t *out_base,
                           void **out_local = nullptr)
{
    if (!e) return ALAC_B200_PARAM_ERROR;
    e->err.clear();
    if (!valid_cfg(cfg) || (!pcm && num_sample_frames) || (!packets_out && !pl) || !packet_sizes) return ALAC_B200_PARAM_ERROR;
    if (pl && (!pl->dst_packets || !pl->exchange || pl->n_ranks == 0 || pl->n_ranks > ALAC_B200_MAX_RANKS || pl->rank >= pl->n_ranks ||
               pl->home_rank >= pl->n_ranks || pl->epoch == 0 || coef_state || (pl->staging && (pl->home_rank != 0 || !pl->slot_offsets))))
        return ALAC_B200_PARAM_ERROR;
    // placement forms: `direct` = the assemble kernel stores at the final offsets (after every rank's scan);
    // `staged` = chunk pipeline, packets leave for the rank's reserved slot while later chunks compute, the home rank compacts
    const bool staged = pl && pl->staging != nullptr, direct = pl && !staged;
    const bool is_home = pl && pl->rank == pl->home_rank;
    const bool to_slot = staged && !is_home;            // this rank's chunks travel to its slot of the home GPU's staging area
    if (out_local) *out_local = nullptr;
    if (out_num_packets) *out_num_packets = 0;
    if (out_bytes) *out_bytes = 0;
    if (out_base) *out_base = 0;
    if (stats) memset(stats, 0, sizeof(*stats));
    CU_CHECK(e, cudaSetDevice(e->device));
    if (e->finish_pending) {        // (an unfinished deferred job)
        CU_CHECK(e, cudaStreamSynchronize(e->side));
        CU_CHECK(e, cudaStreamSynchronize(e->copy_out));
        e->finish_pending = false;
    }
    e->launches = 0;
    e->timers_used = 0;
    e->t_mid.clear();
    std::vector<cudaEvent_t> t_search, t_asm;

    alac_b200_stream whole = {0, num_sample_frames};
    if (!streams) { streams = &whole; n_streams = 1; }
    const uint32_t F = cfg->frame_size, K = cfg->frames_per_segment;
    const uint64_t bpf = (uint64_t)bytes_per_sample(cfg->bit_depth) * cfg->channels;
    // every stream inside the buffer (no wrap-around of first + length)
    for (uint64_t s = 0; s < n_streams; s++)
        if (streams[s].first_sample_frame > num_sample_frames || streams[s].num_sample_frames > num_sample_frames - streams[s].first_sample_frame)
            return ALAC_B200_PARAM_ERROR;

    // ---- packet / segment tables (host).  They depend only on the stream list, the frame size and K, so a call
    //      with the same geometry as the previous one (the usual batch loop) reuses them, on the host and on the device ----
    std::vector<uint64_t> &h_pkt_frame = e->h_pkt_frame;
    std::vector<uint32_t> &h_pkt_samples = e->h_pkt_samples, &h_seg_first = e->h_seg_first, &h_seg_count = e->h_seg_count,
                          &h_seg_stream = e->h_seg_stream;
    const bool same_tables = e->tables_valid && e->tab_F == F && e->tab_K == K && e->tab_streams.size() == n_streams &&
                             memcmp(e->tab_streams.data(), streams, n_streams * sizeof(alac_b200_stream)) == 0;
    if (!same_tables) {
        e->tables_valid = false;
        h_pkt_frame.clear(); h_pkt_samples.clear(); h_seg_first.clear(); h_seg_count.clear(); h_seg_stream.clear();
        e->tab_sample_sum = 0;
        for (uint64_t s = 0; s < n_streams; s++) {
            const alac_b200_stream &st = streams[s];
            const uint64_t packets = (st.num_sample_frames + F - 1) / F;
            if (h_pkt_frame.size() + packets > 0x3fffffffull) return ALAC_B200_PARAM_ERROR;
            const uint32_t first_pkt = (uint32_t)h_pkt_frame.size();
            for (uint64_t p = 0; p < packets; p++) {
                h_pkt_frame.push_back(st.first_sample_frame + p * F);
                h_pkt_samples.push_back((uint32_t)std::min<uint64_t>(F, st.num_sample_frames - p * F));
            }
            e->tab_sample_sum += st.num_sample_frames;
            const uint64_t per_seg = K ? K : std::max<uint64_t>(packets, 1);
            const size_t seg0 = h_seg_first.size();
            for (uint64_t p = 0; p < packets; p += per_seg) {
                h_seg_first.push_back(first_pkt + (uint32_t)p);
                h_seg_count.push_back((uint32_t)std::min<uint64_t>(per_seg, packets - p));
                h_seg_stream.push_back((uint32_t)s);
            }
            if (h_seg_first.size() > seg0) {
                h_seg_stream[seg0] |= 0x80000000u;
                h_seg_stream.back() |= 0x40000000u;
            }
        }
        e->tab_streams.assign(streams, streams + n_streams);
        e->tab_F = F;
        e->tab_K = K;
    }
    const uint32_t P = (uint32_t)h_pkt_frame.size(), S = (uint32_t)h_seg_first.size();
    if (P > sizes_cap) return ALAC_B200_PARAM_ERROR;
    // the exact worst case of THIS packet table (streams may overlap or repeat: every packet is at most its escape
    // form, raw samples + 7 bits per channel of tags + ID_END + padding), not the nominal bound of the buffer length
    const uint64_t need_bytes = e->tab_sample_sum * bpf + (uint64_t)P * (7ull * cfg->channels + 1ull);
    if (!pl && packets_cap < need_bytes) { e->err = "packets_out capacity below the worst case of this packet table"; return ALAC_B200_PARAM_ERROR; }
    if (P == 0 && !pl) return ALAC_B200_OK;

    EncLayout L;
    uint32_t mono_mask, pair_mask;
    build_layout(cfg, L, mono_mask, pair_mask);

    // ---- chunks ----
    // Work is cut into chunks of whole segments.  A chunk bounds the scratch, and when host memory is involved the
    // chunks form a 3-stage pipeline: H2D of chunk c+1 (copy-in stream) overlaps the kernels of chunk c (a compute
    // lane) and the D2H of chunk c-1's packets (copy-out stream).
    const uint32_t cap_words = F + 8;       // >= worst-case Golomb words per channel (<= 32 bits/sample incl. run codes)
    const bool in_host = pcm_mem != ALAC_B200_MEM_DEVICE, out_host = !pl && out_mem != ALAC_B200_MEM_DEVICE;
    const bool sizes_host = out_mem != ALAC_B200_MEM_DEVICE;
    // default: as many packets as fit an 8 GB slab budget (bigger launches fill the GPU more evenly)
    uint64_t chunk_target = scratch_chunk_packets();
    if (!chunk_target) chunk_target = std::max<uint64_t>(4096, (8ull << 30) / ((uint64_t)L.chains_per_packet * cap_words * 4));
    const bool multi = in_host || out_host || staged;   // host buffers / staged placement: chunks run on several compute streams
    if (staged) {
        // Staged placement: a chunk's packets leave for the home GPU when the chunk is done, and whatever has not left
        // when the call's kernels end travels while the caller decodes (defer_finish) -- so few, large chunks (the
        // kernels run at single-stream efficiency: 10-hour 24/96 on two GPUs, 47.9 ms with 13 chunks, 40.4 ms with 4)
        // as long as some bytes start early: four chunks per call, none smaller than 16 k packets.
        chunk_target = std::min<uint64_t>(chunk_target, std::max<uint64_t>(16384, ((uint64_t)P + 3) / 4));
    } else if (multi) {
        chunk_target /= 8;
    }
    const bool taper = multi && !staged && P >= 4096;
    struct Chunk { uint32_t s0, s1, p0, cnt; uint64_t f_lo, f_hi; };
    std::vector<Chunk> chunks;
    uint64_t max_chunk = 0, span_lo = ~0ull, span_hi = 0;
    for (uint32_t s0 = 0; s0 < S;) {
        uint32_t s1 = s0;
        uint64_t cnt = 0;
        uint64_t target = chunk_target;
        if (taper) target = std::min<uint64_t>(target, std::max<uint64_t>(512, tapered_chunk(P, pipeline_chunks(), (uint32_t)std::min<size_t>(chunks.size(), pipeline_chunks() - 1), true)));
        else if (multi) target = std::min<uint64_t>(target, std::max<uint64_t>(2048, (P + pipeline_chunks() - 1) / pipeline_chunks()));
        while (s1 < S && (cnt == 0 || cnt + h_seg_count[s1] <= target)) cnt += h_seg_count[s1++];
        Chunk c;
        c.s0 = s0; c.s1 = s1; c.p0 = h_seg_first[s0]; c.cnt = (uint32_t)cnt;
        c.f_lo = ~0ull; c.f_hi = 0;
        for (uint32_t p = c.p0; p < c.p0 + c.cnt; p++) {
            c.f_lo = std::min(c.f_lo, h_pkt_frame[p]);
            c.f_hi = std::max(c.f_hi, h_pkt_frame[p] + h_pkt_samples[p]);
        }
        span_lo = std::min(span_lo, c.f_lo);
        span_hi = std::max(span_hi, c.f_hi);
        chunks.push_back(c);
        max_chunk = std::max(max_chunk, cnt);
        s0 = s1;
    }
    if (chunks.empty()) { span_lo = span_hi = 0; }
    if (chunks.size() > kMaxChunks) { e->err = "too many chunks"; return ALAC_B200_PARAM_ERROR; }
    // Scratch slots.  Host pipeline: one per compute lane in use (never more lanes than chunks: a frames_per_segment = 0
    // stream is ONE chunk and must not reserve eight whole-stream slabs).  Placed output: assembly waits for the other
    // ranks' totals, i.e. for every chunk's scan, so no slot is reused.  Otherwise chunks run back to back in one slot.
    const uint32_t nstreams_used = multi ? (uint32_t)std::min<size_t>(8, std::max<size_t>(chunks.size(), 1)) : 1u;
    const uint32_t nslots = direct ? (uint32_t)std::max<size_t>(chunks.size(), 1) : nstreams_used;

    // ---- device buffers ----
    CU_CHECK(e, e->pkt_frame.reserve((size_t)P * 8 + 8));
    CU_CHECK(e, e->pkt_samples.reserve((size_t)P * 4 + 4));
    CU_CHECK(e, e->seg_first.reserve((size_t)S * 4 + 4));
    CU_CHECK(e, e->seg_count.reserve((size_t)S * 4 + 4));
    CU_CHECK(e, e->seg_stream.reserve((size_t)S * 4 + 4));
    const size_t recs_per_slot = (size_t)max_chunk * L.elems_per_packet;
    const size_t slab_words_per_slot = (size_t)max_chunk * L.chains_per_packet * cap_words;
    CU_CHECK(e, e->recs.reserve(recs_per_slot * nslots * sizeof(ElemRec) + 64));
    CU_CHECK(e, e->scratch.reserve(slab_words_per_slot * nslots * 4 + 64));
    CU_CHECK(e, e->sizes.reserve((size_t)P * 4 + 4));
    CU_CHECK(e, e->offsets.reserve(((size_t)P + 1) * 8));
    CU_CHECK(e, e->counters.reserve(128));
    CU_CHECK(e, e->xwords.reserve(64));
    CU_CHECK(e, e->scan_tiles.reserve((scan_tiles_for(P) + chunks.size() + 1) * 8));
    // every segment is a single frame and no state is handed over: search and final pass run as two kernels
    const bool split = K == 1 && coef_state == nullptr;
    const size_t jobs_per_slot = (size_t)max_chunk * L.chains_per_packet;
    if (split) {
        CU_CHECK(e, e->jobs.reserve(2 * jobs_per_slot * nslots * sizeof(FinalJob) + 64));
        CU_CHECK(e, e->job_counts.reserve(chunks.size() * 16 + 16));        // per chunk: {4-tap, 8-tap} x {pair, mono launch}
    }
    const uint8_t *d_pcm;
    if (in_host) {
        // only the span of sample-frames the streams touch is staged; d_pcm is biased so kernels index it like the caller's buffer
        CU_CHECK(e, e->pcm.reserve((size_t)((span_hi - span_lo) * bpf) + 64));
        d_pcm = e->pcm.as<uint8_t>() - span_lo * bpf;
    } else {
        d_pcm = static_cast<const uint8_t *>(pcm);
    }
    uint8_t *d_out;
    if (direct || (staged && is_home)) {
        d_out = static_cast<uint8_t *>(pl->dst_packets);        // (the home rank of a staged job owns offset 0: no detour)
    } else if (out_host || to_slot) {
        CU_CHECK(e, e->out.reserve((size_t)need_bytes + 64));
        d_out = e->out.as<uint8_t>();
    } else {
        d_out = static_cast<uint8_t *>(packets_out);
    }

    cudaStream_t st = e->stream;
    unsigned long long h_counters[4] = {0, 0, 0, 0};        // (locals the copy-out stream writes: declared before the guard)
    unsigned long long h_place[3] = {0, 0, 0};
    DrainGuard guard(e);
    CU_CHECK(e, cudaEventRecord(e->ev[0], st));
    if (staged) CU_CHECK(e, cudaEventRecord(e->ev_far0, st));
    // small tables first: they share the H2D copy engine with the PCM chunks and must not queue behind them
    if (!same_tables && P) {
        CU_CHECK(e, cudaMemcpyAsync(e->pkt_frame.p, h_pkt_frame.data(), (size_t)P * 8, cudaMemcpyHostToDevice, st));
        CU_CHECK(e, cudaMemcpyAsync(e->pkt_samples.p, h_pkt_samples.data(), (size_t)P * 4, cudaMemcpyHostToDevice, st));
        CU_CHECK(e, cudaMemcpyAsync(e->seg_first.p, h_seg_first.data(), (size_t)S * 4, cudaMemcpyHostToDevice, st));
        CU_CHECK(e, cudaMemcpyAsync(e->seg_count.p, h_seg_count.data(), (size_t)S * 4, cudaMemcpyHostToDevice, st));
        CU_CHECK(e, cudaMemcpyAsync(e->seg_stream.p, h_seg_stream.data(), (size_t)S * 4, cudaMemcpyHostToDevice, st));
        e->tables_valid = true;     // (a failed call below leaves them valid: they only describe the geometry)
    }
    int16_t *d_state = nullptr;
    if (coef_state) {
        CU_CHECK(e, e->state.reserve((size_t)n_streams * ALAC_B200_STATE_INT16S * 2));
        CU_CHECK(e, cudaMemcpyAsync(e->state.p, coef_state, (size_t)n_streams * ALAC_B200_STATE_INT16S * 2, cudaMemcpyHostToDevice, st));
        d_state = e->state.as<int16_t>();
    }
    CU_CHECK(e, cudaMemsetAsync(e->counters.p, 0, 128, st));
    if (pl) CU_CHECK(e, cudaMemsetAsync(e->xwords.p, 0, 64, st));
    if (P == 0) CU_CHECK(e, cudaMemsetAsync(e->offsets.p, 0, 8, st));       // a rank with no packets still publishes a total
    if (split) CU_CHECK(e, cudaMemsetAsync(e->job_counts.p, 0, chunks.size() * 16 + 16, st));
    // copy-in stream: PCM chunks, each followed by an event the compute stream waits on
    std::vector<cudaEvent_t> h2d_done;
    if (in_host) {
        CU_CHECK(e, cudaStreamWaitEvent(e->copy_in, e->ev[0], 0));
        for (const Chunk &c : chunks) {
            CU_CHECK(e, cudaMemcpyAsync(const_cast<uint8_t *>(d_pcm) + c.f_lo * bpf, static_cast<const uint8_t *>(pcm) + c.f_lo * bpf,
                                        (size_t)((c.f_hi - c.f_lo) * bpf), cudaMemcpyHostToDevice, e->copy_in));
            h2d_done.push_back(e->event_on(e->copy_in));
        }
    }
    CU_CHECK(e, cudaEventRecord(e->ev[1], st));

    // ---- kernels, chunk by chunk.  Device-resident calls stay on the caller's stream; with host buffers chunk c
    //      runs on compute lane c % lanes (own scratch), so its kernels overlap its neighbours' kernels and copies ----
    // coefficients move by at most 1 per predictor step and a frame runs < 2 * frame_size steps on a row:
    // starting from init_coefs (|a| <= 1216) the int16 range cannot be left within K frames if this holds
    const bool wrap = coef_state != nullptr || K == 0 || (uint64_t)K * 2u * F + 1216u > 32767u;
    const bool packed = cfg->channels == 2 && (reinterpret_cast<uintptr_t>(d_pcm) & 7u) == 0;
    // dense: a mono or stereo stream whose packets all start on 16-byte boundaries (see QuadRing, enc_final2_kernel)
    bool dense = cfg->channels <= 2 && (reinterpret_cast<uintptr_t>(d_pcm) & 15u) == 0 && (((uint64_t)F * bpf) % 16 == 0 || P == n_streams);
    for (uint64_t s = 0; s < n_streams && dense; s++) dense = (streams[s].first_sample_frame * bpf) % 16 == 0;
    // developer override: 0 = the generic final pass, 1 = dense one-warp form, 2 = dense two-warp form, default = by job count
    static const int final2_mode = [] { const char *v = getenv("ALAC_B200_FINAL2"); return v ? atoi(v) : -1; }();
    if (final2_mode == 0) dense = false;
    unsigned long long *d_escapes = e->counters.as<unsigned long long>();
    uint32_t *d_max = reinterpret_cast<uint32_t *>(e->counters.as<uint8_t>() + 8);
    // placement words live in their own buffer: a deferred finish (side stream) may still write them while the engine's
    // next call (a decode) reuses the counters
    uint64_t *d_base = e->xwords.as<uint64_t>();            // placed: where this rank's block starts
    uint64_t *d_job_total = e->xwords.as<uint64_t>() + 1;   // placed, home rank: bytes of the whole job
    uint32_t *d_xerr = reinterpret_cast<uint32_t *>(e->xwords.as<uint64_t>() + 2);
    std::vector<cudaEvent_t> comp_done, scan_done;
    int last_dense_form = 0;
    const bool trace = getenv("ALAC_B200_TRACE") != nullptr;
    const auto host_t0 = std::chrono::steady_clock::now();
    size_t tile_at = 0;
    auto assemble_chunk = [&](size_t ci, cudaStream_t cs) {
        const Chunk &c = chunks[ci];
        const uint32_t slot = (uint32_t)(ci % nslots);
        AsmArgs B;
        B.pcm = d_pcm;
        B.pkt_frame = e->pkt_frame.as<uint64_t>();
        B.pkt_samples = e->pkt_samples.as<uint32_t>();
        B.recs = e->recs.as<ElemRec>() + recs_per_slot * slot;
        B.scratch = e->scratch.as<uint32_t>() + slab_words_per_slot * slot;
        B.cap_words = cap_words;
        B.sizes = e->sizes.as<uint32_t>();
        B.offsets = e->offsets.as<uint64_t>();
        B.out = d_out;
        B.base = direct ? d_base : nullptr;
        B.pkt_base = c.p0;
        B.num_packets = c.cnt;
        B.lay = L;
        e->cur = cs;
        t_asm.push_back(e->timer());
        switch (cfg->bit_depth) {
        case 16: enc_launch_assemble<16>(cs, B); break;
        case 20: enc_launch_assemble<20>(cs, B); break;
        case 24: enc_launch_assemble<24>(cs, B); break;
        default: enc_launch_assemble<32>(cs, B); break;
        }
        e->launches++;
        t_asm.push_back(e->timer());
    };
    for (size_t ci = 0; ci < chunks.size(); ci++) {
        const Chunk &c = chunks[ci];
        const uint32_t lane = (uint32_t)(ci % nstreams_used), slot = (uint32_t)(ci % nslots);
        cudaStream_t cs = multi ? e->lanes[lane] : st;
        e->cur = cs;
        if (trace) fprintf(stderr, "[alac_b200] host submits enc chunk %zu at +%.2f ms\n", ci,
                           std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count());
        if (multi && ci < nstreams_used) CU_CHECK(e, cudaStreamWaitEvent(cs, e->ev[1], 0));       // tables are in
        if (in_host) CU_CHECK(e, cudaStreamWaitEvent(cs, h2d_done[ci], 0));
        EncArgs A;
        A.pcm = d_pcm;
        A.pkt_frame = e->pkt_frame.as<uint64_t>();
        A.pkt_samples = e->pkt_samples.as<uint32_t>();
        A.seg_first = e->seg_first.as<uint32_t>();
        A.seg_count = e->seg_count.as<uint32_t>();
        A.seg_stream = e->seg_stream.as<uint32_t>();
        A.seg_base = c.s0;
        A.num_segments = c.s1 - c.s0;
        A.pkt_base = c.p0;
        A.lay = L;
        A.recs = e->recs.as<ElemRec>() + recs_per_slot * slot;
        A.scratch = e->scratch.as<uint32_t>() + slab_words_per_slot * slot;
        A.cap_words = cap_words;
        A.state = d_state;
        JobLists Q;
        Q.max_jobs = (uint32_t)jobs_per_slot;
        Q.jobs = split ? e->jobs.as<FinalJob>() + 2 * jobs_per_slot * slot : nullptr;
        Q.counts = split ? e->job_counts.as<uint32_t>() + 4 * ci : nullptr;
        // two-warp final pass while the launch's jobs do not fill the GPU (about one wave of one-warp CTAs: 148 SMs x 24 x 32 jobs)
        // (chunks of a host-buffer / staged call run side by side on the lanes: what counts is what is on the GPU at once)
        const uint64_t jobs_in_launch = std::min<uint64_t>(P, (uint64_t)c.cnt * nstreams_used) * L.chains_per_packet;
        const int dense_form = !dense ? 0 : final2_mode > 0 ? final2_mode : (jobs_in_launch <= 148ull * 24 * 32 ? 2 : 1);
        last_dense_form = split ? dense_form : 0;
        t_search.push_back(e->timer());
        // (before, between search and final, after) events of the pair launch and of the mono launch (split form)
        cudaEvent_t mid[6];
        for (auto &m : mid) m = e->new_event();
        const JobLists *Qp = split ? &Q : nullptr;
        switch (cfg->bit_depth) {
        case 16: e->launches += enc_launch_search<16>(cs, A, mono_mask, pair_mask, packed, wrap, Qp, mid, dense_form); break;
        case 20: e->launches += enc_launch_search<20>(cs, A, mono_mask, pair_mask, packed, wrap, Qp, mid, dense_form); break;
        case 24: e->launches += enc_launch_search<24>(cs, A, mono_mask, pair_mask, packed, wrap, Qp, mid, dense_form); break;
        default: e->launches += enc_launch_search<32>(cs, A, mono_mask, pair_mask, packed, wrap, Qp, mid, dense_form); break;
        }
        if (split) {
            if (pair_mask) for (int i = 0; i < 3; i++) e->t_mid.push_back(mid[i]);
            if (mono_mask) for (int i = 3; i < 6; i++) e->t_mid.push_back(mid[i]);
        }
        t_search.push_back(e->timer());
        enc_size_kernel<<<(c.cnt + 255) / 256, 256, 0, cs>>>(A.recs, L, cfg->bit_depth, A.pkt_samples + c.p0, c.cnt,
                                                            e->sizes.as<uint32_t>() + c.p0, d_escapes);
        e->launches++;
        // the scan continues from the previous chunk's total: wait for that chunk's scan
        if (multi && ci > 0) CU_CHECK(e, cudaStreamWaitEvent(cs, scan_done[ci - 1], 0));
        launch_scan(e, cs, e->sizes.as<uint32_t>() + c.p0, e->offsets.as<uint64_t>() + c.p0, c.cnt, e->scan_tiles.as<uint64_t>() + tile_at,
                    d_max, ci == 0 ? 0 : 1, &e->h_totals[ci]);
        tile_at += scan_tiles_for(c.cnt);
        if (multi) scan_done.push_back(e->event_on(cs));
        if (!direct) {
            assemble_chunk(ci, cs);
            // running byte total after this chunk -> pinned host word; the host needs it to size the chunk's D2H
            comp_done.push_back(e->event_on(cs));
        }
    }
    Exchange *x = pl ? static_cast<Exchange *>(pl->exchange) : nullptr;
    if (direct) {
        // ---- the cross-GPU step: publish this rank's byte total, wait (on the device) for the ranks in front, then
        //      every chunk's packets go straight to their final place in the destination GPU's buffer
        if (multi) for (cudaEvent_t ev : scan_done) CU_CHECK(e, cudaStreamWaitEvent(st, ev, 0));
        xchg_publish_resolve_kernel<<<1, 32, 0, st>>>(x, pl->rank, pl->epoch, e->offsets.as<uint64_t>() + P, d_base, d_xerr);
        e->launches++;
        for (size_t ci = 0; ci < chunks.size(); ci++) assemble_chunk(ci, st);
        if (pl->dst_sizes && P)
            CU_CHECK(e, cudaMemcpyAsync(pl->dst_sizes + pl->first_packet, e->sizes.p, (size_t)P * 4, cudaMemcpyDefault, st));
        xchg_done_kernel<<<1, 1, 0, st>>>(x, pl->rank, pl->epoch);
        e->launches++;
        if (pl->rank == pl->home_rank) {
            xchg_wait_all_kernel<<<1, 32, 0, st>>>(x, pl->n_ranks, pl->epoch, d_job_total, d_xerr);
            e->launches++;
        }
        comp_done.push_back(e->event_on(st));
    }
    e->cur = nullptr;
    CU_CHECK(e, cudaGetLastError());
    if (multi && !direct) for (cudaEvent_t ev : comp_done) CU_CHECK(e, cudaStreamWaitEvent(st, ev, 0));     // join the lanes
    CU_CHECK(e, cudaEventRecord(e->ev[2], st));

    // ---- results (copy-out stream) ----
    uint64_t total = 0, copied = 0;
    uint8_t *far_out = out_host ? static_cast<uint8_t *>(packets_out)
                                : to_slot ? static_cast<uint8_t *>(pl->staging) + pl->slot_offsets[pl->rank] : nullptr;
    if (to_slot && pl->epoch > 1) {
        // this rank's slot must have been emptied by the home rank's compaction of the previous epoch
        xchg_wait_released_kernel<<<1, 1, 0, e->copy_out>>>(x, pl->epoch - 1u, d_xerr);
        e->launches++;
    }
    for (size_t ci = 0; ci < comp_done.size(); ci++) {
        CU_CHECK(e, cudaEventSynchronize(comp_done[ci]));       // all later GPU work is already queued
        if (direct) break;
        total = e->h_totals[ci];
        if (far_out && total > copied) {
            // host output: D2H; staged placement: a peer copy into this rank's slot on the home GPU (NVLink), while the
            // lanes work on the later chunks
            CU_CHECK(e, cudaMemcpyAsync(far_out + copied, d_out + copied, (size_t)(total - copied), cudaMemcpyDefault, e->copy_out));
            copied = total;
        }
    }
    if (pl && !chunks.empty()) total = e->h_totals[chunks.size() - 1];
    if (staged) {
        // the exchange closes the call: this rank's total, the offset of its block, and "my slot is complete";
        // the home rank then waits for every rank and closes the gaps between the slots inside its own memory
        cudaStream_t xs = is_home ? st : e->copy_out;
        if (pl->dst_sizes && P) CU_CHECK(e, cudaMemcpyAsync(pl->dst_sizes + pl->first_packet, e->sizes.p, (size_t)P * 4, cudaMemcpyDefault, xs));
        xchg_publish_resolve_kernel<<<1, 32, 0, xs>>>(x, pl->rank, pl->epoch, e->offsets.as<uint64_t>() + P, d_base, d_xerr);
        xchg_done_kernel<<<1, 1, 0, xs>>>(x, pl->rank, pl->epoch);
        e->launches += 2;
        if (!is_home) cudaEventRecord(e->ev_far, xs);
        if (is_home) {
            // wait for every rank, close the gaps between the slots, release the slots -- all on the device, on a side
            // stream, so that with defer_finish the host returns as soon as this rank's own block is placed
            CU_CHECK(e, cudaEventRecord(e->ev[2], xs));
            CU_CHECK(e, cudaStreamWaitEvent(e->side, e->ev[2], 0));
            CU_CHECK(e, cudaStreamWaitEvent(e->copy_out, e->ev[2], 0));         // (the result words are read on the copy-out stream)
            SlotOffsets so;
            for (uint32_t r = 0; r < 16; r++) so.v[r] = r < pl->n_ranks ? pl->slot_offsets[r] : 0ull;
            xchg_wait_all_kernel<<<1, 32, 0, e->side>>>(x, pl->n_ranks, pl->epoch, d_job_total, d_xerr);
            static const int compact_ctas = [] { const char *v = getenv("ALAC_B200_COMPACT_CTAS"); return v ? atoi(v) : 148 * 8; }();
            xchg_compact_kernel<<<compact_ctas, 256, 0, e->side>>>(x, pl->n_ranks, pl->epoch, so, static_cast<const uint8_t *>(pl->staging),
                                                             static_cast<uint8_t *>(pl->dst_packets), pl->dst_capacity, d_xerr);
            xchg_release_kernel<<<1, 1, 0, e->side>>>(x, pl->epoch);
            cudaEventRecord(e->ev_far, e->side);
            e->launches += 3;
            e->finish_pending = true;
            if (!pl->defer_finish) {
                CU_CHECK(e, cudaStreamSynchronize(e->side));
                e->finish_pending = false;
                CU_CHECK(e, cudaEventRecord(e->ev[2], xs));      // the job's buffer is complete: end of the kernel phase
            }
        }
    }
    // staged form, a rank other than the home rank, defer_finish: the call ends with this rank's kernels.  Its chunks keep
    // travelling to the home GPU on the copy-out stream (and the exchange words follow them) while the caller already uses
    // the GPU for something else -- decoding its own block; alac_b200_placed_finish() ends the job on this rank too.
    const bool defer_far = to_slot && pl->defer_finish != 0;
    cudaStream_t rs = defer_far ? st : e->copy_out;        // stream of the result copies
    if (P) CU_CHECK(e, cudaMemcpyAsync(packet_sizes, e->sizes.p, (size_t)P * 4, sizes_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, rs));
    CU_CHECK(e, cudaMemcpyAsync(h_counters, e->counters.p, 16, cudaMemcpyDeviceToHost, rs));
    if (pl && !defer_far) CU_CHECK(e, cudaMemcpyAsync(h_place, d_base, 24, cudaMemcpyDeviceToHost, e->copy_out));
    if (coef_state) {
        CU_CHECK(e, cudaMemcpyAsync(coef_state, e->state.p, (size_t)n_streams * ALAC_B200_STATE_INT16S * 2, cudaMemcpyDeviceToHost, e->copy_out));
    }
    CU_CHECK(e, cudaEventRecord(e->ev[3], rs));
    if (!defer_far) CU_CHECK(e, cudaStreamSynchronize(e->copy_out));
    CU_CHECK(e, cudaStreamSynchronize(st));
    guard.armed = false;
    if (defer_far) {
        e->finish_pending = true;
        e->finish_far = true;
        e->finish_total = total;
        e->finish_capacity = pl->dst_capacity;
        if (out_local) *out_local = static_cast<void *>(d_out);
    } else if (pl) {
        e->finish_far = false;
        e->last_base = h_place[0];
        if (h_place[2] & 0xffffffffull) { e->err = "cross-GPU exchange timed out (a rank of the job did not arrive)"; return ALAC_B200_CUDA_ERROR; }
        if (h_place[0] + total > pl->dst_capacity) { e->err = "dst_packets capacity exceeded"; return ALAC_B200_PARAM_ERROR; }
        if (out_base) *out_base = h_place[0];
        if (out_local) *out_local = to_slot ? static_cast<void *>(d_out) : static_cast<void *>(static_cast<uint8_t *>(pl->dst_packets) + h_place[0]);
    }

    if (getenv("ALAC_B200_TRACE") && !direct) {        // developer aid: per-chunk timeline in ms since the call started
        for (size_t ci = 0; ci < chunks.size(); ci++) {
            float a = 0, b = 0, c = 0, d = 0;
            if (in_host) cudaEventElapsedTime(&a, e->ev[0], h2d_done[ci]);
            if (2 * ci + 1 < t_search.size()) { cudaEventElapsedTime(&b, e->ev[0], t_search[2 * ci]); cudaEventElapsedTime(&c, e->ev[0], t_search[2 * ci + 1]); }
            cudaEventElapsedTime(&d, e->ev[0], comp_done[ci]);
            fprintf(stderr, "[alac_b200] enc chunk %zu: %u packets, h2d done %.2f, search %.2f..%.2f, done %.2f\n", ci, chunks[ci].cnt, a, b, c, d);
        }
    }
    if (out_num_packets) *out_num_packets = P;
    if (out_bytes) *out_bytes = total;
    if (stats) {
        stats->num_packets = P;
        stats->payload_bytes = total;
        stats->escape_elements = h_counters[0];
        stats->max_packet_bytes = (uint32_t)(h_counters[1] & 0xffffffffu);
        stats->kernel_launches = e->launches;
        // with host buffers the three phases overlap: h2d = start .. last PCM chunk landed, kernels = first .. last
        // kernel, d2h = last kernel .. last byte on the host
        if (in_host && !h2d_done.empty()) cudaEventElapsedTime(&stats->ms_h2d, e->ev[0], h2d_done.back());
        else cudaEventElapsedTime(&stats->ms_h2d, e->ev[0], e->ev[1]);
        cudaEventElapsedTime(&stats->ms_kernels, e->ev[1], e->ev[2]);
        cudaEventElapsedTime(&stats->ms_d2h, e->ev[2], e->ev[3]);
        for (size_t i = 0; i + 1 < t_search.size(); i += 2) { float ms = 0; cudaEventElapsedTime(&ms, t_search[i], t_search[i + 1]); stats->ms_search += ms; }
        for (size_t i = 0; i + 1 < t_asm.size(); i += 2) { float ms = 0; cudaEventElapsedTime(&ms, t_asm[i], t_asm[i + 1]); stats->ms_assemble += ms; }
        for (size_t i = 0; i + 2 < e->t_mid.size(); i += 3) { float ms = 0; cudaEventElapsedTime(&ms, e->t_mid[i + 1], e->t_mid[i + 2]); stats->ms_final += ms; }
        stats->ms_search -= stats->ms_final;
        stats->final_form = (uint32_t)last_dense_form;
        stats->search_dense = (last_dense_form && ((pair_mask && packed && cfg->bit_depth != 16) || (!pair_mask && mono_mask))) ? 1u : 0u;
        if (pl) stats->payload_bytes = total;
    }
    return ALAC_B200_OK;
}

// ---- several GPUs in one process -------------------------------------------------------------------------------
struct Shard {
    std::vector<alac_b200_stream> streams;
    uint64_t first_packet = 0, packets = 0, first_stream = 0;
    int32_t rc = 0;
    uint64_t np = 0, bytes = 0, base = 0, frames = 0;
    alac_b200_stats st = {};
};

// Contiguous frame ranges, cut at multiples of frames_per_segment packets (whole streams when a stream is one chain)
static std::vector<Shard> plan_encode_shards(const alac_b200_enc_config *cfg, const alac_b200_stream *streams, uint64_t n_streams,
                                             uint32_t n_dev, bool whole_streams)
{
    const uint64_t F = cfg->frame_size, K = cfg->frames_per_segment;
    uint64_t P = 0;
    for (uint64_t s = 0; s < n_streams; s++) P += (streams[s].num_sample_frames + F - 1) / F;
    const uint64_t per_dev = std::max<uint64_t>(1, (P + n_dev - 1) / n_dev);
    std::vector<Shard> sh(n_dev);
    std::vector<bool> seen(n_dev, false);
    uint64_t cum = 0;
    for (uint64_t s = 0; s < n_streams; s++) {
        const uint64_t ps = (streams[s].num_sample_frames + F - 1) / F;
        for (uint64_t p = 0; p < ps;) {
            const uint32_t d = (uint32_t)std::min<uint64_t>(n_dev - 1, cum / per_dev);
            const uint64_t room = (uint64_t)(d + 1) * per_dev - cum;
            uint64_t take = ps - p;
            if (!whole_streams && K && d + 1 < n_dev) take = std::min<uint64_t>(take, (room + K - 1) / K * K);
            if (!seen[d]) { seen[d] = true; sh[d].first_packet = cum; sh[d].first_stream = s; }
            alac_b200_stream piece;
            piece.first_sample_frame = streams[s].first_sample_frame + p * F;
            piece.num_sample_frames = std::min<uint64_t>(take * F, streams[s].num_sample_frames - p * F);
            sh[d].streams.push_back(piece);
            sh[d].packets += take;
            cum += take;
            p += take;
        }
    }
    while (!sh.empty() && sh.back().packets == 0) sh.pop_back();        // devices without work take no part
    return sh;
}

static void merge_stats(alac_b200_stats *dst, const alac_b200_stats &s)
{
    dst->num_packets += s.num_packets;
    dst->payload_bytes += s.payload_bytes;
    dst->escape_elements += s.escape_elements;
    dst->max_packet_bytes = std::max(dst->max_packet_bytes, s.max_packet_bytes);
    dst->kernel_launches += s.kernel_launches;
    float *a = &dst->ms_h2d;
    const float *b = &s.ms_h2d;
    for (int i = 0; i < 10; i++) a[i] = std::max(a[i], b[i]);         // the ten ms_* fields: slowest device
}

static int32_t multi_encode(alac_b200_engine *e, const alac_b200_enc_config *cfg, const void *pcm, uint64_t num_sample_frames,
                            int32_t pcm_mem, const alac_b200_stream *streams, uint64_t n_streams, void *packets_out,
                            uint64_t packets_cap, uint32_t *packet_sizes, uint64_t sizes_cap, int32_t out_mem, int16_t *coef_state,
                            uint64_t *out_num_packets, uint64_t *out_bytes, alac_b200_stats *stats)
{
    e->err.clear();
    if (!valid_cfg(cfg) || (!pcm && num_sample_frames) || !packets_out || !packet_sizes) return ALAC_B200_PARAM_ERROR;
    if (out_num_packets) *out_num_packets = 0;
    if (out_bytes) *out_bytes = 0;
    if (stats) memset(stats, 0, sizeof(*stats));
    alac_b200_stream whole = {0, num_sample_frames};
    if (!streams) { streams = &whole; n_streams = 1; }
    for (uint64_t s = 0; s < n_streams; s++)
        if (streams[s].first_sample_frame > num_sample_frames || streams[s].num_sample_frames > num_sample_frames - streams[s].first_sample_frame)
            return ALAC_B200_PARAM_ERROR;
    const bool whole_streams = cfg->frames_per_segment == 0 || coef_state != nullptr;
    std::vector<Shard> sh = plan_encode_shards(cfg, streams, n_streams, (uint32_t)e->subs.size(), whole_streams);
    uint64_t P = 0, sample_sum = 0;
    for (const Shard &h : sh) P += h.packets;
    for (uint64_t s = 0; s < n_streams; s++) sample_sum += streams[s].num_sample_frames;
    if (P > sizes_cap) return ALAC_B200_PARAM_ERROR;
    const uint64_t bpf = (uint64_t)bytes_per_sample(cfg->bit_depth) * cfg->channels;
    if (packets_cap < sample_sum * bpf + P * (7ull * cfg->channels + 1ull)) { e->err = "packets_out capacity below the worst case of this packet table"; return ALAC_B200_PARAM_ERROR; }
    if (P == 0) return ALAC_B200_OK;
    const bool out_dev = out_mem == ALAC_B200_MEM_DEVICE && coef_state == nullptr;    // (a state hand-off takes the block-copy form)
    const uint32_t epoch = ++e->epoch;
    const uint32_t m = (uint32_t)sh.size();
    // device output: every GPU's assemble kernel stores straight into packets_out (home device) at its final offset;
    // host output: every GPU keeps its block, the blocks go down over each GPU's own PCIe link once the totals are known
    std::vector<std::thread> th;
    for (uint32_t d = 0; d < m; d++) {
        th.emplace_back([&, d] {
            alac_b200_engine *se = e->subs[d];
            Shard &h = sh[d];
            memset(&h.st, 0, sizeof(h.st));
            int16_t *state_d = coef_state ? coef_state + h.first_stream * ALAC_B200_STATE_INT16S : nullptr;
            if (out_dev) {
                alac_b200_placement pl;
                pl.dst_packets = packets_out; pl.dst_capacity = packets_cap; pl.dst_sizes = nullptr; pl.first_packet = h.first_packet;
                pl.exchange = e->xchg.p; pl.rank = d; pl.n_ranks = m; pl.home_rank = 0; pl.epoch = epoch;
                pl.staging = nullptr; pl.slot_offsets = nullptr; pl.defer_finish = 0;
                h.rc = encode_core(se, cfg, pcm, num_sample_frames, pcm_mem, h.streams.data(), h.streams.size(), nullptr, 0,
                                   packet_sizes + h.first_packet, h.packets, ALAC_B200_MEM_DEVICE, nullptr, &h.np, &h.bytes, &h.st, &pl, &h.base);
            } else {
                uint64_t need = 0;
                for (const alac_b200_stream &q : h.streams) need += q.num_sample_frames;
                need = need * bpf + h.packets * (7ull * cfg->channels + 1ull);
                cudaSetDevice(se->device);
                if (se->m_out.reserve((size_t)need + 64) != cudaSuccess || se->m_sizes.reserve((size_t)h.packets * 4 + 4) != cudaSuccess) { h.rc = ALAC_B200_MEM_ERROR; return; }
                h.rc = encode_core(se, cfg, pcm, num_sample_frames, pcm_mem, h.streams.data(), h.streams.size(), se->m_out.p, need,
                                   se->m_sizes.as<uint32_t>(), h.packets, ALAC_B200_MEM_DEVICE, state_d, &h.np, &h.bytes, &h.st, nullptr, nullptr);
            }
        });
    }
    for (auto &t : th) t.join();
    int32_t rc = 0;
    for (uint32_t d = 0; d < m && !rc; d++)
        if (sh[d].rc) { rc = sh[d].rc; e->err = "device " + std::to_string(e->subs[d]->device) + ": " + e->subs[d]->err; }
    uint64_t total = 0;
    if (!rc && !out_dev) {
        for (uint32_t d = 0; d < m; d++) {
            alac_b200_engine *se = e->subs[d];
            cudaSetDevice(se->device);
            if (cudaMemcpyAsync(static_cast<uint8_t *>(packets_out) + total, se->m_out.p, (size_t)sh[d].bytes, cudaMemcpyDefault, se->stream) != cudaSuccess ||
                cudaMemcpyAsync(packet_sizes + sh[d].first_packet, se->m_sizes.p, (size_t)sh[d].packets * 4, cudaMemcpyDefault, se->stream) != cudaSuccess)
                rc = ALAC_B200_CUDA_ERROR;
            total += sh[d].bytes;
        }
        for (uint32_t d = 0; d < m; d++) { cudaSetDevice(e->subs[d]->device); if (cudaStreamSynchronize(e->subs[d]->stream) != cudaSuccess) rc = ALAC_B200_CUDA_ERROR; }
    } else {
        for (uint32_t d = 0; d < m; d++) total += sh[d].bytes;
    }
    cudaSetDevice(e->device);
    if (rc) return rc;
    if (out_num_packets) *out_num_packets = P;
    if (out_bytes) *out_bytes = total;
    if (stats) for (uint32_t d = 0; d < m; d++) merge_stats(stats, sh[d].st);
    return ALAC_B200_OK;
}

extern "C" int32_t alac_b200_encode(alac_b200_engine *e, const alac_b200_enc_config *cfg, const void *pcm,
                                    uint64_t num_sample_frames, int32_t pcm_mem, const alac_b200_stream *streams,
                                    uint64_t n_streams, void *packets_out, uint64_t packets_cap, uint32_t *packet_sizes,
                                    uint64_t sizes_cap, int32_t out_mem, int16_t *coef_state, uint64_t *out_num_packets,
                                    uint64_t *out_bytes, alac_b200_stats *stats)
{
    if (!e) return ALAC_B200_PARAM_ERROR;
    if (e->subs.size() > 1)
        return multi_encode(e, cfg, pcm, num_sample_frames, pcm_mem, streams, n_streams, packets_out, packets_cap, packet_sizes, sizes_cap,
                            out_mem, coef_state, out_num_packets, out_bytes, stats);
    return encode_core(e, cfg, pcm, num_sample_frames, pcm_mem, streams, n_streams, packets_out, packets_cap, packet_sizes, sizes_cap,
                       out_mem, coef_state, out_num_packets, out_bytes, stats, nullptr, nullptr);
}

extern "C" int32_t alac_b200_encode_placed(alac_b200_engine *e, const alac_b200_enc_config *cfg, const void *pcm,
                                           uint64_t num_sample_frames, int32_t pcm_mem, const alac_b200_stream *streams,
                                           uint64_t n_streams, const alac_b200_placement *placement, uint32_t *packet_sizes,
                                           uint64_t sizes_cap, int32_t out_mem, uint64_t *out_num_packets, uint64_t *out_bytes,
                                           uint64_t *out_base, void **out_local_block, alac_b200_stats *stats)
{
    if (!e || !placement || e->subs.size() > 1) return ALAC_B200_PARAM_ERROR;
    return encode_core(e, cfg, pcm, num_sample_frames, pcm_mem, streams, n_streams, nullptr, 0, packet_sizes, sizes_cap, out_mem, nullptr,
                       out_num_packets, out_bytes, stats, placement, out_base, out_local_block);
}

extern "C" int32_t alac_b200_placed_finish(alac_b200_engine *e, uint64_t *out_job_bytes)
{
    if (!e) return ALAC_B200_PARAM_ERROR;
    if (out_job_bytes) *out_job_bytes = 0;
    if (!e->finish_pending) return ALAC_B200_OK;
    CU_CHECK(e, cudaSetDevice(e->device));
    CU_CHECK(e, cudaStreamSynchronize(e->finish_far ? e->copy_out : e->side));
    e->finish_pending = false;
    if (getenv("ALAC_B200_TRACE")) {        // developer aid: when the cross-GPU work of the job ended, in ms since the encode call started
        float ms = 0;
        cudaEventElapsedTime(&ms, e->ev_far0, e->ev_far);
        fprintf(stderr, "[alac_b200] dev %d: %s at %.2f ms after the placed call started\n", e->device,
                e->finish_far ? "last block byte handed to the home GPU" : "gaps closed, job buffer complete", ms);
    }
    unsigned long long h[3] = {0, 0, 0};
    CU_CHECK(e, cudaMemcpy(h, e->xwords.p, 24, cudaMemcpyDeviceToHost));
    if (h[2] & 0xffffffffull) {
        e->err = (h[2] & 0xffffffffull) == 3 ? "dst_packets capacity exceeded" : "cross-GPU exchange timed out (a rank of the job did not arrive)";
        return (h[2] & 0xffffffffull) == 3 ? ALAC_B200_PARAM_ERROR : ALAC_B200_CUDA_ERROR;
    }
    if (e->finish_far) {
        e->last_base = h[0];
        if (h[0] + e->finish_total > e->finish_capacity) { e->err = "dst_packets capacity exceeded"; return ALAC_B200_PARAM_ERROR; }
    } else if (out_job_bytes) {
        *out_job_bytes = h[1];
    }
    return ALAC_B200_OK;
}

extern "C" int32_t alac_b200_placed_base(alac_b200_engine *e, uint64_t *out_base)
{
    if (!e || !out_base || e->finish_pending) return ALAC_B200_PARAM_ERROR;
    *out_base = e->last_base;
    return ALAC_B200_OK;
}

extern "C" int32_t alac_b200_engine_create_multi(const int32_t *devices, uint32_t n_devices, alac_b200_engine **out_engine)
{
    if (!out_engine) return ALAC_B200_PARAM_ERROR;
    *out_engine = nullptr;
    if (!devices || n_devices == 0 || n_devices > ALAC_B200_MAX_RANKS) return ALAC_B200_PARAM_ERROR;
    for (uint32_t i = 0; i < n_devices; i++)
        for (uint32_t j = 0; j < i; j++)
            if (devices[i] == devices[j] || devices[i] < 0) return ALAC_B200_PARAM_ERROR;
    alac_b200_engine *home = nullptr;
    int32_t rc = alac_b200_engine_create(devices[0], &home);
    if (rc) return rc;
    home->subs.push_back(home);
    for (uint32_t i = 1; i < n_devices && !rc; i++) {
        alac_b200_engine *se = nullptr;
        rc = alac_b200_engine_create(devices[i], &se);
        if (!rc) home->subs.push_back(se);
    }
    // peer access between every pair: kernels of any device store into the home device's buffers (and may read PCM
    // that lives on another device)
    for (uint32_t i = 0; i < n_devices && !rc; i++) {
        for (uint32_t j = 0; j < n_devices && !rc; j++) {
            if (i == j) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) != cudaSuccess || !can) { rc = ALAC_B200_UNIMPLEMENTED; break; }
            cudaSetDevice(devices[i]);
            const cudaError_t pe = cudaDeviceEnablePeerAccess(devices[j], 0);
            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) rc = ALAC_B200_CUDA_ERROR;
            cudaGetLastError();
        }
    }
    cudaSetDevice(devices[0]);
    if (!rc && (home->xchg.reserve(ALAC_B200_EXCHANGE_BYTES) != cudaSuccess ||
                cudaMemset(home->xchg.p, 0, ALAC_B200_EXCHANGE_BYTES) != cudaSuccess))
        rc = ALAC_B200_CUDA_ERROR;
    if (rc) { alac_b200_engine_destroy(home); return rc; }
    *out_engine = home;
    return ALAC_B200_OK;
}

extern "C" uint32_t alac_b200_engine_num_devices(const alac_b200_engine *e) { return e ? (uint32_t)std::max<size_t>(1, e->subs.size()) : 0u; }

// plain cudaMalloc memory (exportable with cudaIpcGetMemHandle) and the IPC wrappers of a one-process-per-GPU launcher
extern "C" int32_t alac_b200_device_alloc(alac_b200_engine *e, uint64_t bytes, void **out_ptr)
{
    if (!e || !out_ptr) return ALAC_B200_PARAM_ERROR;
    *out_ptr = nullptr;
    CU_CHECK(e, cudaSetDevice(e->device));
    CU_CHECK(e, cudaMalloc(out_ptr, (size_t)std::max<uint64_t>(bytes, 1)));
    return ALAC_B200_OK;
}
extern "C" int32_t alac_b200_device_free(alac_b200_engine *e, void *ptr)
{
    if (!e) return ALAC_B200_PARAM_ERROR;
    CU_CHECK(e, cudaSetDevice(e->device));
    CU_CHECK(e, cudaFree(ptr));
    return ALAC_B200_OK;
}
extern "C" int32_t alac_b200_ipc_export(alac_b200_engine *e, void *ptr, void *out_handle64)
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the ABI");
    if (!e || !ptr || !out_handle64) return ALAC_B200_PARAM_ERROR;
    CU_CHECK(e, cudaSetDevice(e->device));
    cudaIpcMemHandle_t h;
    CU_CHECK(e, cudaIpcGetMemHandle(&h, ptr));
    memcpy(out_handle64, &h, 64);
    return ALAC_B200_OK;
}
extern "C" int32_t alac_b200_ipc_open(alac_b200_engine *e, const void *handle64, void **out_ptr)
{
    if (!e || !handle64 || !out_ptr) return ALAC_B200_PARAM_ERROR;
    *out_ptr = nullptr;
    CU_CHECK(e, cudaSetDevice(e->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CU_CHECK(e, cudaIpcOpenMemHandle(out_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return ALAC_B200_OK;
}
extern "C" int32_t alac_b200_ipc_close(alac_b200_engine *e, void *ptr)
{
    if (!e || !ptr) return ALAC_B200_PARAM_ERROR;
    CU_CHECK(e, cudaSetDevice(e->device));
    CU_CHECK(e, cudaIpcCloseMemHandle(ptr));
    return ALAC_B200_OK;
}

// ------------------------------------------------------------------------------------------------
// decode
// ------------------------------------------------------------------------------------------------
static void configure_decode_kernels()
{
    dec_configure<16>(); dec_configure<20>(); dec_configure<24>(); dec_configure<32>();
}

static int32_t decode_core(alac_b200_engine *e, const void *cookie, uint32_t cookie_size, const void *packets,
                           const uint32_t *packet_sizes, uint64_t num_packets, int32_t in_mem, void *pcm_out,
                           uint64_t pcm_cap, uint32_t *packet_samples, int32_t *packet_status, int32_t out_mem,
                           uint64_t *out_sample_frames, alac_b200_stats *stats)
{
    if (!e) return ALAC_B200_PARAM_ERROR;
    e->err.clear();
    if (out_sample_frames) *out_sample_frames = 0;
    if (stats) memset(stats, 0, sizeof(*stats));
    uint32_t f[11];
    int32_t rc = alac_b200_parse_cookie(cookie, cookie_size, f);
    if (rc) return rc;
    const uint32_t frame_length = f[0], depth = f[2], nch = f[6];
    if (!valid_depth(depth) || nch == 0 || frame_length == 0) return ALAC_B200_PARAM_ERROR;
    if (num_packets > 0x3fffffffull) return ALAC_B200_PARAM_ERROR;
    if ((!packets || !packet_sizes || !pcm_out) && num_packets) return ALAC_B200_PARAM_ERROR;
    if (num_packets == 0) return ALAC_B200_OK;
    CU_CHECK(e, cudaSetDevice(e->device));
    if (!e->decode_configured) { configure_decode_kernels(); e->decode_configured = true; }
    e->launches = 0;
    e->timers_used = 0;
    e->t_mid.clear();
    const uint32_t P = (uint32_t)num_packets;
    const uint64_t bpf = (uint64_t)bytes_per_sample(depth) * nch;
    cudaStream_t st = e->stream;
    const bool in_host = in_mem != ALAC_B200_MEM_DEVICE, out_host = out_mem != ALAC_B200_MEM_DEVICE;

    // ---- chunks: with host output the packets are processed in up to 8 pipelined chunks (H2D of chunk c+1,
    //      kernels of chunk c and D2H of chunk c-1's PCM overlap); with device output one chunk and a capacity check
    struct Chunk { uint32_t p0, cnt; uint64_t b0, b1; };
    std::vector<Chunk> chunks;
    uint64_t total_bytes = 0;
    {
        const bool taper = out_host && P >= 4096;
        for (uint32_t p0 = 0; p0 < P;) {
            uint32_t per = P;
            if (taper) per = (uint32_t)std::max<uint64_t>(512, tapered_chunk(P, pipeline_chunks(), (uint32_t)std::min<size_t>(chunks.size(), pipeline_chunks() - 1), false));
            else if (out_host) per = std::max<uint32_t>(2048, (P + pipeline_chunks() - 1) / pipeline_chunks());     // see multi below
            Chunk c;
            c.p0 = p0; c.cnt = std::min(per, P - p0); c.b0 = total_bytes;
            if (in_host) for (uint32_t i = p0; i < p0 + c.cnt; i++) total_bytes += packet_sizes[i];
            c.b1 = total_bytes;
            chunks.push_back(c);
            p0 += c.cnt;
        }
    }
    if (chunks.size() > kMaxChunks) { e->err = "too many chunks"; return ALAC_B200_PARAM_ERROR; }
    uint32_t max_cnt = 0;
    for (const Chunk &c : chunks) max_cnt = std::max(max_cnt, c.cnt);
    const uint32_t groups = (max_cnt + 31) / 32;
    const bool multi = out_host;                        // chunks run on several compute streams
    const uint32_t nlanes = multi ? (uint32_t)std::min<size_t>(8, chunks.size()) : 1u;     // never more scratch slots than chunks
    const size_t chan_words_per_lane = (size_t)groups * 32 * nch * frame_length;

    CU_CHECK(e, e->d_pkt_off.reserve(((size_t)P + 1) * 8));
    CU_CHECK(e, e->d_pkt_samples.reserve((size_t)P * 4));
    CU_CHECK(e, e->d_out_frame.reserve(((size_t)P + 1) * 8));
    CU_CHECK(e, e->d_status.reserve((size_t)P * 4));
    CU_CHECK(e, e->d_class.reserve((size_t)P * 4));
    CU_CHECK(e, e->d_rank.reserve((size_t)P * 4));
    CU_CHECK(e, e->d_perm.reserve((size_t)P * 4));
    CU_CHECK(e, e->d_chan.reserve(chan_words_per_lane * nlanes * 4));
    CU_CHECK(e, e->d_meta.reserve((size_t)P * nch * sizeof(DecChanMeta)));
    CU_CHECK(e, e->d_hdr.reserve((size_t)P * nch * sizeof(DecChanHdr)));
    CU_CHECK(e, e->counters.reserve(64 * chunks.size()));
    CU_CHECK(e, e->scan_tiles.reserve((2 * scan_tiles_for(P) + chunks.size() + 2) * 8));

    // per-packet status on the host: pinned (a copy into pageable memory blocks the caller until everything queued before
    // it on its stream has run -- on a rank of a staged job that is the block still travelling to the home GPU)
    if (e->h_status_cap < P) {
        if (e->h_status) cudaFreeHost(e->h_status);
        e->h_status = nullptr; e->h_status_cap = 0;
        const size_t want = (size_t)P + P / 8 + 64;
        CU_CHECK(e, cudaHostAlloc(&e->h_status, want * sizeof(int32_t), cudaHostAllocDefault));
        e->h_status_cap = want;
    }
    int32_t *h_status = e->h_status;
    DrainGuard guard(e);
    const bool trace_host = getenv("ALAC_B200_TRACE") != nullptr;
    const auto host_t0 = std::chrono::steady_clock::now();
    auto host_ms = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count(); };
    double hp[5] = {0, 0, 0, 0, 0};
    CU_CHECK(e, cudaEventRecord(e->ev[0], st));
    const uint32_t *d_sizes;
    const uint8_t *d_packets;
    std::vector<cudaEvent_t> h2d_done;
    if (in_host) {
        CU_CHECK(e, e->d_sizes.reserve((size_t)P * 4));
        CU_CHECK(e, e->d_packets.reserve((size_t)total_bytes + 64));
        CU_CHECK(e, cudaStreamWaitEvent(e->copy_in, e->ev[0], 0));
        CU_CHECK(e, cudaMemcpyAsync(e->d_sizes.p, packet_sizes, (size_t)P * 4, cudaMemcpyHostToDevice, e->copy_in));
        CU_CHECK(e, cudaEventRecord(e->ev[3], e->copy_in));
        CU_CHECK(e, cudaStreamWaitEvent(st, e->ev[3], 0));                 // sizes are in
        for (const Chunk &c : chunks) {
            if (c.b1 > c.b0)
                CU_CHECK(e, cudaMemcpyAsync(e->d_packets.as<uint8_t>() + c.b0, static_cast<const uint8_t *>(packets) + c.b0,
                                            (size_t)(c.b1 - c.b0), cudaMemcpyHostToDevice, e->copy_in));
            h2d_done.push_back(e->event_on(e->copy_in));
        }
        d_sizes = e->d_sizes.as<uint32_t>();
        d_packets = e->d_packets.as<uint8_t>();
    } else {
        d_sizes = packet_sizes;
        d_packets = static_cast<const uint8_t *>(packets);
    }
    uint8_t *d_pcm;
    if (out_host) {
        CU_CHECK(e, e->d_pcm.reserve((size_t)((uint64_t)P * frame_length * bpf) + 64));
        d_pcm = e->d_pcm.as<uint8_t>();
    } else {
        d_pcm = static_cast<uint8_t *>(pcm_out);
    }
    CU_CHECK(e, cudaMemsetAsync(e->counters.p, 0, 64 * chunks.size(), st));     // class counters of every chunk
    // byte offsets of all packets at once (the sizes are all known up front)
    launch_scan(e, st, d_sizes, e->d_pkt_off.as<uint64_t>(), P, e->scan_tiles.as<uint64_t>(), nullptr, 0, nullptr);
    size_t tile_at = scan_tiles_for(P);
    CU_CHECK(e, cudaEventRecord(e->ev[1], st));

    DecArgs A;
    A.packets = d_packets;
    A.pkt_off = e->d_pkt_off.as<uint64_t>();
    A.pkt_size = d_sizes;
    A.frame_length = frame_length;
    A.pb = f[3];
    A.mb = f[4];
    A.kb = f[5];
    A.num_channels = nch;
    A.pcm_out = d_pcm;
    A.out_frame = e->d_out_frame.as<uint64_t>();
    A.pkt_samples = e->d_pkt_samples.as<uint32_t>();
    A.pkt_status = e->d_status.as<int32_t>();
    A.pkt_class = e->d_class.as<uint32_t>();
    A.pkt_rank = e->d_rank.as<uint32_t>();
    A.class_count = e->counters.as<uint32_t>();
    A.perm = e->d_perm.as<uint32_t>();
    A.chan_scratch = e->d_chan.as<int32_t>();
    A.chan_meta = e->d_meta.as<DecChanMeta>();
    A.chan_hdr = e->d_hdr.as<DecChanHdr>();

    std::vector<cudaEvent_t> comp_done, scan_done, t_dec;
    uint64_t total_frames = 0;
    for (size_t ci = 0; ci < chunks.size(); ci++) {
        const Chunk &c = chunks[ci];
        const uint32_t lane = (uint32_t)(ci % nlanes);
        cudaStream_t cs = multi ? e->lanes[lane] : st;
        e->cur = cs;
        if (multi && ci < nlanes) CU_CHECK(e, cudaStreamWaitEvent(cs, e->ev[1], 0));       // packet offsets are in
        if (in_host) CU_CHECK(e, cudaStreamWaitEvent(cs, h2d_done[ci], 0));
        A.pkt_base = c.p0;
        A.num_packets = c.cnt;
        A.class_count = e->counters.as<uint32_t>() + 16 * ci;
        A.chan_scratch = e->d_chan.as<int32_t>() + chan_words_per_lane * lane;
        dec_header_kernel<<<(c.cnt + 127) / 128, 128, 0, cs>>>(A);
        dec_perm_kernel<<<(c.cnt + 127) / 128, 128, 0, cs>>>(A);
        // output positions continue from the previous chunk's total
        if (multi && ci > 0) CU_CHECK(e, cudaStreamWaitEvent(cs, scan_done[ci - 1], 0));
        launch_scan(e, cs, A.pkt_samples + c.p0, e->d_out_frame.as<uint64_t>() + c.p0, c.cnt, e->scan_tiles.as<uint64_t>() + tile_at, nullptr,
                    ci == 0 ? 0 : 1, &e->h_totals[ci]);
        tile_at += scan_tiles_for(c.cnt);
        if (multi) scan_done.push_back(e->event_on(cs));
        e->launches += 2;
        if (!out_host && (uint64_t)P * frame_length * bpf > pcm_cap) {
            // a caller-owned device buffer that could not hold P full packets: the sample counts must be known to fit
            // before anything is written (a buffer sized for full packets needs no check and no host round trip)
            CU_CHECK(e, cudaStreamSynchronize(cs));
            if (e->h_totals[ci] * bpf > pcm_cap) { e->err = "pcm capacity exceeded"; return ALAC_B200_PARAM_ERROR; }
        }
        t_dec.push_back(e->timer());
        // regular mono / stereo groups go to the fused kernel (entropy and finish warps side by side).  Not for the depths with
        // shift bytes (24 / 32 bit): there the output phase merges bytes re-read from the packet, too much for the single
        // finish warp of the fused kernel (one hour of 24/96: 6.8 ms either way; ten hours: 59.6 ms fused against 54.3 ms)
        static const int fused_mode = [] { const char *v = getenv("ALAC_B200_FUSED"); return v ? atoi(v) : -1; }();   // developer override: 0 never, 1 always
        A.fused = (nch <= 2 && (fused_mode < 0 ? (depth == 16 || depth == 20) : fused_mode != 0)) ? 1u : 0u;
        cudaEvent_t mid[2] = {e->new_event(), e->new_event()};
        switch (depth) {
        case 16: e->launches += dec_launch_main<16>(cs, A, mid); break;
        case 20: e->launches += dec_launch_main<20>(cs, A, mid); break;
        case 24: e->launches += dec_launch_main<24>(cs, A, mid); break;
        default: e->launches += dec_launch_main<32>(cs, A, mid); break;
        }
        e->t_mid.push_back(mid[0]);
        e->t_mid.push_back(mid[1]);
        t_dec.push_back(e->timer());
        comp_done.push_back(e->event_on(cs));
    }
    e->cur = nullptr;
    CU_CHECK(e, cudaGetLastError());
    if (multi) for (cudaEvent_t ev : comp_done) CU_CHECK(e, cudaStreamWaitEvent(st, ev, 0));     // join the lanes
    CU_CHECK(e, cudaEventRecord(e->ev[2], st));
    hp[0] = host_ms();

    // ---- results (copy-out stream) ----
    uint64_t copied = 0;
    bool overflow = false;
    for (size_t ci = 0; ci < chunks.size(); ci++) {
        CU_CHECK(e, cudaEventSynchronize(comp_done[ci]));
        total_frames = e->h_totals[ci];
        if (out_host) {
            uint64_t upto = total_frames * bpf;
            if (upto > pcm_cap) { upto = pcm_cap; overflow = true; }
            if (upto > copied) {
                CU_CHECK(e, cudaMemcpyAsync(static_cast<uint8_t *>(pcm_out) + copied, d_pcm + copied, (size_t)(upto - copied),
                                            cudaMemcpyDeviceToHost, e->copy_out));
                copied = upto;
            }
        }
    }
    hp[1] = host_ms();
    // device-resident output: the small result copies follow the kernels on the call's own stream (the copy-out stream
    // may still be moving a deferred placed block: nothing of this call has to wait for that)
    cudaStream_t rs = out_host ? e->copy_out : st;
    CU_CHECK(e, cudaMemcpyAsync(h_status, A.pkt_status, (size_t)P * 4, cudaMemcpyDeviceToHost, rs));
    hp[2] = host_ms();
    const cudaMemcpyKind to_user = out_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (packet_samples) CU_CHECK(e, cudaMemcpyAsync(packet_samples, A.pkt_samples, (size_t)P * 4, to_user, rs));
    if (packet_status) CU_CHECK(e, cudaMemcpyAsync(packet_status, A.pkt_status, (size_t)P * 4, to_user, rs));
    CU_CHECK(e, cudaEventRecord(e->ev[3], rs));
    if (out_host) CU_CHECK(e, cudaStreamSynchronize(e->copy_out));
    hp[3] = host_ms();
    CU_CHECK(e, cudaStreamSynchronize(st));
    hp[4] = host_ms();
    guard.armed = false;
    if (trace_host) fprintf(stderr, "[alac_b200] dev %d decode host: queued %.2f, kernels done %.2f, status copy issued %.2f, copy-out drained %.2f, end %.2f ms\n",
                            e->device, hp[0], hp[1], hp[2], hp[3], hp[4]);
    if (overflow) { e->err = "pcm capacity exceeded"; return ALAC_B200_PARAM_ERROR; }

    if (getenv("ALAC_B200_TRACE")) {        // developer aid: per-chunk timeline in ms since the call started
        for (size_t ci = 0; ci < chunks.size(); ci++) {
            float a = 0, b = 0, c = 0, d = 0;
            if (in_host) cudaEventElapsedTime(&a, e->ev[0], h2d_done[ci]);
            cudaEventElapsedTime(&b, e->ev[0], t_dec[2 * ci]);
            cudaEventElapsedTime(&c, e->ev[0], t_dec[2 * ci + 1]);
            cudaEventElapsedTime(&d, e->ev[0], comp_done[ci]);
            fprintf(stderr, "[alac_b200] dec chunk %zu: %u packets, h2d done %.2f, kernels %.2f..%.2f, done %.2f\n", ci, chunks[ci].cnt, a, b, c, d);
        }
        float z = 0;
        cudaEventElapsedTime(&z, e->ev[0], e->ev[3]);
        fprintf(stderr, "[alac_b200] dec last byte on the host at %.2f\n", z);
    }
    if (out_sample_frames) *out_sample_frames = total_frames;
    int32_t first_err = 0;
    for (uint32_t i = 0; i < P && !first_err; i++) first_err = h_status[i];
    if (stats) {
        stats->num_packets = P;
        stats->payload_bytes = total_frames * bpf;
        stats->kernel_launches = e->launches;
        if (in_host) cudaEventElapsedTime(&stats->ms_h2d, e->ev[0], h2d_done.back());
        else cudaEventElapsedTime(&stats->ms_h2d, e->ev[0], e->ev[1]);
        cudaEventElapsedTime(&stats->ms_kernels, e->ev[1], e->ev[2]);
        cudaEventElapsedTime(&stats->ms_d2h, e->ev[2], e->ev[3]);
        for (size_t i = 0; i + 1 < t_dec.size(); i += 2) {
            float ms = 0;
            cudaEventElapsedTime(&ms, t_dec[i], t_dec[i + 1]); stats->ms_decode += ms;
            cudaEventElapsedTime(&ms, t_dec[i], e->t_mid[i]); stats->ms_fused += ms;
            cudaEventElapsedTime(&ms, e->t_mid[i], e->t_mid[i + 1]); stats->ms_entropy += ms;
            cudaEventElapsedTime(&ms, e->t_mid[i + 1], t_dec[i + 1]); stats->ms_finish += ms;
        }
    }
    return first_err;
}

// several GPUs: contiguous packet ranges; every GPU decodes its range into its own memory, then the blocks go to their
// place in pcm_out (per-packet sample counts are data: the offsets are only known once every range is decoded)
static int32_t multi_decode(alac_b200_engine *e, const void *cookie, uint32_t cookie_size, const void *packets,
                            const uint32_t *packet_sizes, uint64_t num_packets, int32_t in_mem, void *pcm_out, uint64_t pcm_cap,
                            uint32_t *packet_samples, int32_t *packet_status, int32_t out_mem, uint64_t *out_sample_frames,
                            alac_b200_stats *stats)
{
    e->err.clear();
    if (out_sample_frames) *out_sample_frames = 0;
    if (stats) memset(stats, 0, sizeof(*stats));
    uint32_t f[11];
    int32_t rc = alac_b200_parse_cookie(cookie, cookie_size, f);
    if (rc) return rc;
    const uint32_t frame_length = f[0], depth = f[2], nch = f[6];
    if (!valid_depth(depth) || nch == 0 || frame_length == 0 || num_packets > 0x3fffffffull) return ALAC_B200_PARAM_ERROR;
    if ((!packets || !packet_sizes || !pcm_out) && num_packets) return ALAC_B200_PARAM_ERROR;
    if (num_packets == 0) return ALAC_B200_OK;
    const uint64_t bpf = (uint64_t)bytes_per_sample(depth) * nch;
    // byte offsets of the shard boundaries need the sizes on the host
    std::vector<uint32_t> h_sizes;
    const uint32_t *hs = packet_sizes;
    if (in_mem == ALAC_B200_MEM_DEVICE) {
        h_sizes.resize((size_t)num_packets);
        cudaSetDevice(e->device);
        if (cudaMemcpy(h_sizes.data(), packet_sizes, (size_t)num_packets * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return ALAC_B200_CUDA_ERROR;
        hs = h_sizes.data();
    }
    const uint32_t n_dev = (uint32_t)e->subs.size();
    const uint64_t per_dev = (num_packets + n_dev - 1) / n_dev;
    std::vector<Shard> sh;
    uint64_t at = 0, byte_at = 0;
    std::vector<uint64_t> byte_off;
    for (uint32_t d = 0; d < n_dev && at < num_packets; d++) {
        Shard h;
        h.first_packet = at;
        h.packets = std::min<uint64_t>(per_dev, num_packets - at);
        byte_off.push_back(byte_at);
        for (uint64_t i = at; i < at + h.packets; i++) byte_at += hs[i];
        at += h.packets;
        sh.push_back(h);
    }
    const uint32_t m = (uint32_t)sh.size();
    std::vector<std::thread> th;
    for (uint32_t d = 0; d < m; d++) {
        th.emplace_back([&, d] {
            alac_b200_engine *se = e->subs[d];
            Shard &h = sh[d];
            memset(&h.st, 0, sizeof(h.st));
            cudaSetDevice(se->device);
            const uint64_t cap = h.packets * frame_length * bpf;
            if (se->m_out.reserve((size_t)cap + 64) != cudaSuccess || se->m_aux.reserve((size_t)h.packets * 8 + 8) != cudaSuccess) { h.rc = ALAC_B200_MEM_ERROR; return; }
            uint32_t *d_samples = se->m_aux.as<uint32_t>();
            int32_t *d_status = reinterpret_cast<int32_t *>(d_samples + h.packets);
            const uint32_t *sizes_d = packet_sizes + h.first_packet;
            if (in_mem != ALAC_B200_MEM_DEVICE) {
                // host input: decode_core stages packets and sizes itself
                h.rc = decode_core(se, cookie, cookie_size, static_cast<const uint8_t *>(packets) + byte_off[d], sizes_d, h.packets, ALAC_B200_MEM_HOST,
                                   se->m_out.p, cap, d_samples, d_status, ALAC_B200_MEM_DEVICE, &h.frames, &h.st);
            } else {
                h.rc = decode_core(se, cookie, cookie_size, static_cast<const uint8_t *>(packets) + byte_off[d], sizes_d, h.packets, ALAC_B200_MEM_DEVICE,
                                   se->m_out.p, cap, d_samples, d_status, ALAC_B200_MEM_DEVICE, &h.frames, &h.st);
            }
        });
    }
    for (auto &t : th) t.join();
    int32_t first_err = 0;
    for (uint32_t d = 0; d < m; d++) {
        // a packet status (kALAC_ParamError) is a result, not a failure of the call: the blocks are still delivered
        if (sh[d].rc && sh[d].rc != ALAC_B200_PARAM_ERROR) { cudaSetDevice(e->device); e->err = "device " + std::to_string(e->subs[d]->device) + ": " + e->subs[d]->err; return sh[d].rc; }
        if (sh[d].rc && !e->subs[d]->err.empty()) { cudaSetDevice(e->device); e->err = e->subs[d]->err; return sh[d].rc; }
        if (sh[d].rc && !first_err) first_err = sh[d].rc;
    }
    uint64_t frames = 0;
    bool overflow = false;
    for (uint32_t d = 0; d < m; d++) {
        alac_b200_engine *se = e->subs[d];
        cudaSetDevice(se->device);
        uint64_t nbytes = sh[d].frames * bpf;
        if (frames * bpf + nbytes > pcm_cap) { overflow = true; nbytes = pcm_cap > frames * bpf ? pcm_cap - frames * bpf : 0; }
        cudaError_t ce = cudaSuccess;
        if (nbytes) ce = cudaMemcpyAsync(static_cast<uint8_t *>(pcm_out) + frames * bpf, se->m_out.p, (size_t)nbytes, cudaMemcpyDefault, se->stream);
        uint32_t *d_samples = se->m_aux.as<uint32_t>();
        if (ce == cudaSuccess && packet_samples) ce = cudaMemcpyAsync(packet_samples + sh[d].first_packet, d_samples, (size_t)sh[d].packets * 4, cudaMemcpyDefault, se->stream);
        if (ce == cudaSuccess && packet_status) ce = cudaMemcpyAsync(packet_status + sh[d].first_packet, d_samples + sh[d].packets, (size_t)sh[d].packets * 4, cudaMemcpyDefault, se->stream);
        if (ce != cudaSuccess) { cudaSetDevice(e->device); e->err = cudaGetErrorString(ce); return ALAC_B200_CUDA_ERROR; }
        frames += sh[d].frames;
    }
    for (uint32_t d = 0; d < m; d++) { cudaSetDevice(e->subs[d]->device); cudaStreamSynchronize(e->subs[d]->stream); }
    cudaSetDevice(e->device);
    if (overflow) { e->err = "pcm capacity exceeded"; return ALAC_B200_PARAM_ERROR; }
    if (out_sample_frames) *out_sample_frames = frames;
    if (stats) { for (uint32_t d = 0; d < m; d++) merge_stats(stats, sh[d].st); stats->payload_bytes = frames * bpf; }
    return first_err;
}

extern "C" int32_t alac_b200_decode(alac_b200_engine *e, const void *cookie, uint32_t cookie_size, const void *packets,
                                    const uint32_t *packet_sizes, uint64_t num_packets, int32_t in_mem, void *pcm_out,
                                    uint64_t pcm_cap, uint32_t *packet_samples, int32_t *packet_status, int32_t out_mem,
                                    uint64_t *out_sample_frames, alac_b200_stats *stats)
{
    if (!e) return ALAC_B200_PARAM_ERROR;
    if (e->subs.size() > 1)
        return multi_decode(e, cookie, cookie_size, packets, packet_sizes, num_packets, in_mem, pcm_out, pcm_cap, packet_samples,
                            packet_status, out_mem, out_sample_frames, stats);
    return decode_core(e, cookie, cookie_size, packets, packet_sizes, num_packets, in_mem, pcm_out, pcm_cap, packet_samples,
                       packet_status, out_mem, out_sample_frames, stats);
}

// ---- asynchronous forms: the synchronous call on a worker thread of the engine ------------------------------------
extern "C" int32_t alac_b200_encode_submit(alac_b200_engine *e, const alac_b200_enc_config *cfg, const void *pcm,
                                           uint64_t num_sample_frames, int32_t pcm_mem, const alac_b200_stream *streams,
                                           uint64_t n_streams, void *packets_out, uint64_t packets_cap, uint32_t *packet_sizes,
                                           uint64_t sizes_cap, int32_t out_mem, int16_t *coef_state, uint64_t *out_num_packets,
                                           uint64_t *out_bytes, alac_b200_stats *stats)
{
    if (!e || !cfg || e->async_busy) return ALAC_B200_PARAM_ERROR;
    if (e->worker.joinable()) e->worker.join();
    e->async_cfg = *cfg;
    e->async_streams.clear();
    if (streams) e->async_streams.assign(streams, streams + n_streams);
    e->async_busy = true;
    e->worker = std::thread([=] {
        e->async_rc = alac_b200_encode(e, &e->async_cfg, pcm, num_sample_frames, pcm_mem, streams ? e->async_streams.data() : nullptr, n_streams,
                                       packets_out, packets_cap, packet_sizes, sizes_cap, out_mem, coef_state, out_num_packets, out_bytes, stats);
    });
    return ALAC_B200_OK;
}

extern "C" int32_t alac_b200_decode_submit(alac_b200_engine *e, const void *cookie, uint32_t cookie_size, const void *packets,
                                           const uint32_t *packet_sizes, uint64_t num_packets, int32_t in_mem, void *pcm_out,
                                           uint64_t pcm_cap, uint32_t *packet_samples, int32_t *packet_status, int32_t out_mem,
                                           uint64_t *out_sample_frames, alac_b200_stats *stats)
{
    if (!e || !cookie || e->async_busy) return ALAC_B200_PARAM_ERROR;
    if (e->worker.joinable()) e->worker.join();
    e->async_cookie.assign(static_cast<const uint8_t *>(cookie), static_cast<const uint8_t *>(cookie) + cookie_size);
    e->async_busy = true;
    e->worker = std::thread([=] {
        e->async_rc = alac_b200_decode(e, e->async_cookie.data(), cookie_size, packets, packet_sizes, num_packets, in_mem, pcm_out, pcm_cap,
                                       packet_samples, packet_status, out_mem, out_sample_frames, stats);
    });
    return ALAC_B200_OK;
}

extern "C" int32_t alac_b200_wait(alac_b200_engine *e)
{
    if (!e || !e->async_busy) return ALAC_B200_PARAM_ERROR;
    if (e->worker.joinable()) e->worker.join();
    e->async_busy = false;
    return e->async_rc;
}

// ------------------------------------------------------------------------------------------------
// CAF packet table -> packet sizes, on the device
// ------------------------------------------------------------------------------------------------
extern "C" int32_t alac_b200_ber_table_sizes(alac_b200_engine *e, const void *table, uint64_t table_bytes, int32_t table_mem,
                                             uint64_t data_bytes, uint32_t *sizes_out, uint64_t sizes_cap, int32_t out_mem,
                                             uint64_t *out_num_packets)
{
    if (!e || !out_num_packets || (!table && table_bytes) || (!sizes_out && sizes_cap)) return ALAC_B200_PARAM_ERROR;
    e->err.clear();
    *out_num_packets = 0;
    if (table_bytes == 0) return ALAC_B200_OK;
    if (table_bytes > 0x7fffffffull) return ALAC_B200_PARAM_ERROR;
    CU_CHECK(e, cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    const uint8_t *d_table;
    if (table_mem == ALAC_B200_MEM_DEVICE) {
        d_table = static_cast<const uint8_t *>(table);
    } else {
        CU_CHECK(e, e->d_packets.reserve((size_t)table_bytes + 64));
        CU_CHECK(e, cudaMemcpyAsync(e->d_packets.p, table, (size_t)table_bytes, cudaMemcpyHostToDevice, st));
        d_table = e->d_packets.as<uint8_t>();
    }
    // every entry is at least one byte, so there are at most table_bytes entries
    CU_CHECK(e, e->d_class.reserve((size_t)table_bytes * 4));                 // end-of-entry marks
    CU_CHECK(e, e->d_pkt_off.reserve(((size_t)table_bytes + 1) * 8));         // entry index of every byte
    CU_CHECK(e, e->d_sizes.reserve((size_t)table_bytes * 4));                 // decoded sizes
    CU_CHECK(e, e->d_out_frame.reserve(((size_t)table_bytes + 1) * 8));       // byte offsets of the packets
    CU_CHECK(e, e->counters.reserve(256));
    const uint32_t blocks = (uint32_t)((table_bytes + 255) / 256);
    ber_flag_kernel<<<blocks, 256, 0, st>>>(d_table, table_bytes, e->d_class.as<uint32_t>());
    CU_CHECK(e, e->scan_tiles.reserve(2 * scan_tiles_for(table_bytes) * 8));
    launch_scan(e, st, e->d_class.as<uint32_t>(), e->d_pkt_off.as<uint64_t>(), table_bytes, e->scan_tiles.as<uint64_t>(), nullptr, 0, nullptr);
    CU_CHECK(e, cudaMemsetAsync(e->d_sizes.p, 0, (size_t)table_bytes * 4, st));
    ber_value_kernel<<<blocks, 256, 0, st>>>(d_table, table_bytes, e->d_pkt_off.as<uint64_t>(), e->d_sizes.as<uint32_t>(), table_bytes);
    uint64_t entries = 0;
    CU_CHECK(e, cudaMemcpyAsync(&entries, e->d_pkt_off.as<uint64_t>() + table_bytes, 8, cudaMemcpyDeviceToHost, st));
    CU_CHECK(e, cudaStreamSynchronize(st));
    uint64_t n = entries;
    if (entries) {
        unsigned long long *d_first = e->counters.as<unsigned long long>() + 16;
        const unsigned long long init = entries;
        CU_CHECK(e, cudaMemcpyAsync(d_first, &init, 8, cudaMemcpyHostToDevice, st));
        launch_scan(e, st, e->d_sizes.as<uint32_t>(), e->d_out_frame.as<uint64_t>(), entries, e->scan_tiles.as<uint64_t>() + scan_tiles_for(table_bytes), nullptr, 0, nullptr);
        ber_count_kernel<<<(uint32_t)((entries + 255) / 256), 256, 0, st>>>(e->d_sizes.as<uint32_t>(), e->d_out_frame.as<uint64_t>(), entries,
                                                                          data_bytes, d_first);
        unsigned long long first = 0;
        CU_CHECK(e, cudaMemcpyAsync(&first, d_first, 8, cudaMemcpyDeviceToHost, st));
        CU_CHECK(e, cudaStreamSynchronize(st));
        n = first;
    }
    CU_CHECK(e, cudaGetLastError());
    if (n > sizes_cap) { e->err = "sizes capacity exceeded"; return ALAC_B200_PARAM_ERROR; }
    if (n) {
        CU_CHECK(e, cudaMemcpyAsync(sizes_out, e->d_sizes.p, (size_t)n * 4,
                                    out_mem == ALAC_B200_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
        CU_CHECK(e, cudaStreamSynchronize(st));
    }
    *out_num_packets = n;
    return ALAC_B200_OK;
}

// packet sizes -> BER table (the 'pakt' payload after its 24-byte header), on the device
extern "C" int32_t alac_b200_ber_table_build(alac_b200_engine *e, const uint32_t *sizes, uint64_t num_packets, int32_t sizes_mem,
                                             void *table_out, uint64_t table_cap, int32_t out_mem, uint64_t *out_table_bytes)
{
    if (!e || !out_table_bytes || (!sizes && num_packets) || (!table_out && num_packets)) return ALAC_B200_PARAM_ERROR;
    e->err.clear();
    *out_table_bytes = 0;
    if (num_packets == 0) return ALAC_B200_OK;
    if (num_packets > 0x3fffffffull) return ALAC_B200_PARAM_ERROR;
    CU_CHECK(e, cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    const uint32_t *d_sizes;
    if (sizes_mem == ALAC_B200_MEM_DEVICE) {
        d_sizes = sizes;
    } else {
        CU_CHECK(e, e->d_sizes.reserve((size_t)num_packets * 4));
        CU_CHECK(e, cudaMemcpyAsync(e->d_sizes.p, sizes, (size_t)num_packets * 4, cudaMemcpyHostToDevice, st));
        d_sizes = e->d_sizes.as<uint32_t>();
    }
    CU_CHECK(e, e->d_class.reserve((size_t)num_packets * 4));                 // entry lengths
    CU_CHECK(e, e->d_pkt_off.reserve(((size_t)num_packets + 1) * 8));         // entry offsets
    const uint32_t blocks = (uint32_t)((num_packets + 255) / 256);
    ber_len_kernel<<<blocks, 256, 0, st>>>(d_sizes, num_packets, e->d_class.as<uint32_t>());
    CU_CHECK(e, e->scan_tiles.reserve(scan_tiles_for(num_packets) * 8));
    launch_scan(e, st, e->d_class.as<uint32_t>(), e->d_pkt_off.as<uint64_t>(), num_packets, e->scan_tiles.as<uint64_t>(), nullptr, 0, &e->h_totals[0]);
    CU_CHECK(e, cudaStreamSynchronize(st));
    const uint64_t total = e->h_totals[0];
    if (total > table_cap) { e->err = "table capacity exceeded"; return ALAC_B200_PARAM_ERROR; }
    uint8_t *d_table;
    if (out_mem == ALAC_B200_MEM_DEVICE) {
        d_table = static_cast<uint8_t *>(table_out);
    } else {
        CU_CHECK(e, e->d_packets.reserve((size_t)total + 64));
        d_table = e->d_packets.as<uint8_t>();
    }
    ber_emit_kernel<<<blocks, 256, 0, st>>>(d_sizes, e->d_pkt_off.as<uint64_t>(), num_packets, d_table);
    CU_CHECK(e, cudaGetLastError());
    if (out_mem != ALAC_B200_MEM_DEVICE) CU_CHECK(e, cudaMemcpyAsync(table_out, d_table, (size_t)total, cudaMemcpyDeviceToHost, st));
    CU_CHECK(e, cudaStreamSynchronize(st));
    *out_table_bytes = total;
    return ALAC_B200_OK;
}
