// alac_decode.cuh -- decode kernels.
//
//   dec_header_kernel   one lane per packet: reads the first audio element's header to learn the
//                       packet's sample count (partial-frame field), so output offsets can be scanned.
//                       Also classifies the packet (classes 0..3 = "regular": one compressed element matching the
//                       channel count, mode 0, denShift 9, 4 or 8 taps); dec_perm_kernel turns the classes into a
//                       lane -> packet permutation so warps run uniform tap counts and regular packets fill groups.
//   dec_fused_kernel    groups of 32 regular mono / stereo packets: an entropy warp (lane = packet: dyn_decomp of
//                       both channels into shared-memory tiles) and a finish warp (unpc_block down each lane's
//                       column, then un-mix / pack / store with lane = sample) run side by side, handing tiles
//                       over through named barriers.  The residuals never leave the SM.
//   dec_entropy_kernel  every other group, one lane per packet: the general element loop
//                       (codec/ALACDecoder.cu:571-1002): header parse, dyn_decomp, escape samples, FIL / DSE.
//                       Each channel's residuals go to a scratch laid out
//                       [group of 32 packets][channel][sample][lane], so every store of a warp is one
//                       128-byte line; a DecChanMeta / DecChanHdr per channel says how to finish it.
//   dec_finish_kernel   the same groups, one lane per (packet, element): unpc_block, then the data-parallel tail
//                       (codec/ALACDecoder.cu:193-495 unmixNN / copyPredictorToNN).  Residual tiles of 32 samples x
//                       32 packets stream through shared memory (cp.async, one tile ahead); the predictor runs down
//                       each lane's column in place, then the warp flips roles (lane = sample) to un-mix, merge
//                       the shift bytes read straight from the packet, pack and store each packet's 32
//                       sample-frames as one contiguous run.
//   ber_*_kernel        CAF 'pakt' table <-> packet sizes on the device.
// The Golomb streams are read through BitReader (alac_device.cuh): a branch-free 64-bit window over a cp.async ring.
#pragma once
#include "alac_device.cuh"

namespace alacb {

struct DecArgs {
    const uint8_t *packets;
    const uint64_t *pkt_off;      // byte offset of each packet (exclusive scan of sizes)
    const uint32_t *pkt_size;
    uint32_t pkt_base;            // chunk: first packet of this launch (perm / scratch slots are chunk-relative)
    uint32_t num_packets;         // chunk: packets in this launch
    uint32_t frame_length, pb, mb, kb, num_channels;
    uint8_t *pcm_out;
    const uint64_t *out_frame;    // first output sample-frame of each packet (exclusive scan)
    uint32_t *pkt_samples;
    int32_t *pkt_status;
    // lane -> packet permutation that groups packets by predictor order (warp-uniform tap counts)
    uint32_t *pkt_class, *pkt_rank, *class_count, *perm;
    int32_t *chan_scratch;        // [group][channel][frame_length][32]
    struct DecChanMeta *chan_meta; // [packet][channel]
    struct DecChanHdr *chan_hdr;   // [packet][channel]
    uint32_t fused;                // dec_fused_kernel takes the regular groups of this call
};

// how dec_finish_kernel finishes one channel of one packet
enum : uint32_t { CH_ZERO = 0, CH_MONO = 1, CH_PAIR_U = 2, CH_PAIR_V = 3 };
struct DecChanMeta {
    uint32_t n;             // samples
    uint32_t shift_pos;     // bit position of the element's shift region inside the packet
    uint8_t kind;           // CH_*
    uint8_t shift;          // shifted-off bits per sample (0, 8, 16)
    int8_t mix_res;
    uint8_t mix_bits;
};

constexpr uint32_t kDecClasses = 16;
constexpr uint32_t kRegularClasses = 4;     // classes below this are taken by dec_fused_kernel

// walk element tags until the first SCE/LFE/CPE and return its sample count
static __global__ void dec_header_kernel(DecArgs A)
{
    const uint32_t local = blockIdx.x * blockDim.x + threadIdx.x;
    if (local >= A.num_packets) return;
    const uint32_t p = A.pkt_base + local;
    const uint32_t size = A.pkt_size[p];
    BitPeek br;
    br.start(A.packets + A.pkt_off[p], size);
    uint32_t n = A.frame_length;
    uint32_t cls = kDecClasses - 1;     // escape / no audio element / unusual orders
    for (int guard = 0; guard < 64; guard++) {
        if (!((br.pos >> 3) < size)) break;
        const uint32_t tag = br.get(3);
        if (tag == ID_SCE || tag == ID_LFE || tag == ID_CPE) {
            if (tag == ID_CPE && A.num_channels < 2) break;    // a pair that does not fit ends the packet unread (:759-760)
            br.pos += 4;
            const bool hdr_ok = br.get(12) == 0;
            const uint32_t hb = br.get(4);
            if (hb >> 3) { n = br.get(16) << 16; n |= br.get(16); }
            const bool first = true;        // skipped FIL / DSE elements in front do not matter
            if (!(hb & 1u) && ((hb >> 1) & 3u) != 3u) {
                // predictor set-up of the first element.  Classes 0..3 are "regular" (what dec_fused_kernel takes):
                // the element matches the channel count, mode 0, denShift 9, 4 or 8 taps -> class = 2 * [U has 8] + [V has 8].
                // Other compressed elements: class = 4 + 3 * order(U) + order(V), order in {4, 8, other}.
                const bool pair = (tag == ID_CPE);
                br.pos += 16;
                const uint32_t mu = br.get(8), nu = br.get(8) & 0x1fu;
                uint32_t mv = (0u << 4) | kDenShift, nv = 4;
                if (pair) { br.pos += nu * 16; mv = br.get(8); nv = br.get(8) & 0x1fu; }
                const bool layout_ok = first && hdr_ok && (A.num_channels == (pair ? 2u : 1u)) && n <= A.frame_length;
                const bool fast = mu == kDenShift && mv == kDenShift && (nu == 4 || nu == 8) && (nv == 4 || nv == 8);
                if (layout_ok && fast) cls = 2 * (nu == 8 ? 1u : 0u) + (nv == 8 ? 1u : 0u);
                else cls = kRegularClasses + 3 * (nu == 4 ? 0u : nu == 8 ? 1u : 2u) + (nv == 4 ? 0u : nv == 8 ? 1u : 2u);
            }
            break;
        } else if (tag == ID_DSE) {                 // codec/ALACDecoder.cu:1033-1059
            br.pos += 4;
            const uint32_t align = br.get(1);
            uint32_t count = br.get(8);
            if (count == 255) count += br.get(8);
            if (align && (br.pos & 7u)) br.pos += 8u - (br.pos & 7u);
            br.pos += count * 8;
        } else if (tag == ID_FIL) {                 // codec/ALACDecoder.cu:1012-1027
            int32_t count = (int32_t)br.get(4);
            if (count == 15) count += (int32_t)br.get(8) - 1;
            br.pos += (uint32_t)count * 8;
        } else {
            break;
        }
    }
    A.pkt_samples[p] = n <= A.frame_length ? n : 0u;
    A.pkt_class[p] = cls;
    A.pkt_rank[p] = atomicAdd(&A.class_count[cls], 1u);
}

// perm[first slot of the packet's class + its rank inside the class] = packet
static __global__ void dec_perm_kernel(DecArgs A)
{
    const uint32_t local = blockIdx.x * blockDim.x + threadIdx.x;
    if (local >= A.num_packets) return;
    const uint32_t p = A.pkt_base + local;
    const uint32_t cls = A.pkt_class[p];
    uint32_t first = 0;
    for (uint32_t c = 0; c < cls; c++) first += A.class_count[c];
    A.perm[A.pkt_base + first + A.pkt_rank[p]] = p;
}

// what the entropy kernel hands to the predictor kernel for one channel of one packet
struct DecChanHdr {
    uint8_t mode, den_shift, num, chan_bits;    // num = kRawChannel: samples are final (escape element), no predictor
    int16_t coefs[32];
};
constexpr uint32_t kRawChannel = 0xffu;

struct ChanHeader {
    uint32_t mode, den_shift, pb_factor, num;
    int16_t coefs[32];
};

__device__ __forceinline__ void read_chan_header(BitPeek &br, ChanHeader &h)
{
    uint32_t hb = br.get(8);                        // codec/ALACDecoder.cu:660-669
    h.mode = hb >> 4;
    h.den_shift = hb & 0xfu;
    hb = br.get(8);
    h.pb_factor = hb >> 5;
    h.num = hb & 0x1fu;
    for (uint32_t i = 0; i < h.num; i++) h.coefs[i] = (int16_t)br.get(16);
}

__device__ __forceinline__ void store_chan_hdr(DecChanHdr *dst, const ChanHeader &h, uint32_t chan_bits)
{
    dst->mode = (uint8_t)h.mode;
    dst->den_shift = (uint8_t)h.den_shift;
    dst->num = (uint8_t)h.num;
    dst->chan_bits = (uint8_t)chan_bits;
    for (uint32_t i = 0; i < h.num; i++) dst->coefs[i] = h.coefs[i];
}

// dyn_decomp for one channel (codec/ag_dec.c:272-362): n residuals -> this lane's column of the channel tile
__device__ __forceinline__ int32_t entropy_channel(BitReader &br, BitPeek &bp, uint32_t cap_bits, const DecArgs &A, uint32_t n,
                                                   uint32_t chan_bits, const ChanHeader &h, int32_t *dst)
{
    br.seek(bp.pos);
    AgDec ag;
    ag.start(br, n, A.mb, (A.pb * h.pb_factor) / 4, A.kb, chan_bits);       // codec/ALACDecoder.cu:682
    for (uint32_t j = 0; j < n; j++) {
        dst[(size_t)j * 32u] = ag.at(br, j);
        if ((j & (kTopUpEvery - 1u)) == kTopUpEvery - 1u) br.top_up();
    }
    ag.finish(br, cap_bits);
    bp.pos = br.pos;
    return ag.status;
}

constexpr uint32_t kTileRows = 32;
constexpr uint32_t kTilePitch = 36;     // words per row: 32 lanes + 4 pad keeps rows 16-byte aligned and the
                                        // transposed read of phase 3 at 4-way bank conflicts at most
constexpr uint32_t kTileWords = kTileRows * kTilePitch;


// what the parallel phase needs to know about one packet; the first 16 bytes are all the common case reads
struct __align__(16) FinMeta {
    uint8_t *out0;          // where sample-frame 0 of this packet's channel slot goes
    uint32_t n;
    uint8_t kind, shift, mix_bits;
    int8_t mix_res;
    const uint8_t *pkt;     // for the shift region
    uint32_t pkt_size, shift_pos;
};
static_assert(sizeof(FinMeta) == 32, "FinMeta is read as two 16-byte halves");

// 16-byte cp.async with zero-fill (src_bytes = 0 reads nothing)
__device__ __forceinline__ void cp_async_16(uint32_t smem_dst, const void *gsrc, uint32_t src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}

__device__ __forceinline__ uint4 lds_meta(uint32_t smem_addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_addr) : "memory");
    return v;
}

// un-mix / merge / store one tile: lane = sample j0 + lane, loop over the group's packets (codec/ALACDecoder.cu:193-495).
// `metas` is the shared-memory ADDRESS of the FinMeta array (taken once by the caller: inside the loop the compiler
// re-derived it from the generic pointer with an S2R at every trip).  `simple16` (warp-uniform) = 16-bit output, every
// packet of the group a stereo pair without shift bytes and the output 4-byte aligned: a short branch-free body.
template <int DEPTH>
__device__ __forceinline__ void flush_tile(const DecArgs &A, uint32_t metas, const int32_t *bu, const int32_t *bv, uint32_t j0,
                                           uint32_t lane, uint32_t pr0, uint32_t pr_step, bool out_pair32, uint32_t zero_chans,
                                           bool simple16)
{
    constexpr uint32_t bps = DepthTraits<DEPTH>::kBytes;
    const uint32_t stride = A.num_channels * bps;
    const uint32_t j = j0 + lane;
    if (DEPTH == 16 && simple16) {
        for (uint32_t pr = pr0; pr < 32; pr += pr_step) {
            const uint4 q = lds_meta(metas + pr * (uint32_t)sizeof(FinMeta));   // out0, n, {kind, shift, mix_bits, mix_res}
            if (j >= q.z) continue;                                             // (slots past the last packet have n = 0)
            const int32_t l = bu[lane * kTilePitch + pr], v = bv[lane * kTilePitch + pr];
            const int32_t mix_res = (int32_t)q.w >> 24;
            const int32_t lm = l + v - ((mix_res * v) >> ((q.w >> 16) & 31u));  // :193-223
            const int32_t lo = mix_res ? lm : l, ro = mix_res ? lm - v : v;
            uint8_t *out = reinterpret_cast<uint8_t *>(((uint64_t)q.y << 32) | q.x) + (size_t)j * 4u;
            *reinterpret_cast<uint32_t *>(out) = ((uint32_t)lo & 0xffffu) | ((uint32_t)ro << 16);
        }
        return;
    }
    // Packets in batches of kBatch: the batch's shift-region words (two aligned loads per sample: addresses come from the
    // packets' metadata, so nothing can be requested earlier) are all requested before the first one is used -- one
    // round trip to L2 / HBM per batch instead of one per packet.
    constexpr uint32_t kBatch = 4;
    for (uint32_t prb = pr0; prb < 32; prb += kBatch * pr_step) {
        uint4 qs[kBatch];
        uint32_t w0[kBatch], w1[kBatch], sh_bit[kBatch], sh_mode[kBatch];      // sh_mode: 0 none, 1 the two words, 2 bounds-checked reader
#pragma unroll
        for (uint32_t u = 0; u < kBatch; u++) {
            const uint32_t pr = prb + u * pr_step;
            sh_mode[u] = 0; w0[u] = w1[u] = 0; sh_bit[u] = 0;
            qs[u] = make_uint4(0u, 0u, 0u, (uint32_t)CH_PAIR_V);                 // (skipped below)
            if (pr >= 32) continue;
            const uint4 q = lds_meta(metas + pr * (uint32_t)sizeof(FinMeta));   // out0, n, {kind, shift, mix_bits, mix_res}
            qs[u] = q;
            const uint32_t n = q.z, kind = q.w & 0xffu, shift = (q.w >> 8) & 0xffu;
            if (kind == CH_PAIR_V || kind == CH_ZERO || j >= n || !shift) continue;
            // this sample's shifted-off low bits (mono: `shift` bits, pair: L then R, 2 * shift bits) straight from the packet
            const uint4 q2 = lds_meta(metas + pr * (uint32_t)sizeof(FinMeta) + 16u);    // pkt, pkt_size, shift_pos
            const uint64_t addr = ((uint64_t)q2.y << 32) | q2.x;
            const uint32_t W = (kind == CH_MONO) ? shift : 2u * shift;                  // <= 32
            const uint32_t bias = (uint32_t)(addr & 3u) * 8u;
            const uint32_t abs_bit = bias + q2.w + j * W, i = abs_bit >> 5;
            sh_bit[u] = abs_bit & 31u;
            if ((i + 2u) * 32u <= bias + q2.z * 8u) {
                // both words lie inside the packet: two aligned loads and a funnel shift
                const uint32_t *base = reinterpret_cast<const uint32_t *>(addr & ~(uint64_t)3);
                w0[u] = __ldg(base + i);
                w1[u] = __ldg(base + i + 1);
                sh_mode[u] = 1;
            } else {
                sh_mode[u] = 2;                                                         // near the end: the bounds-checked reader
            }
        }
#pragma unroll
        for (uint32_t u = 0; u < kBatch; u++) {
            const uint32_t pr = prb + u * pr_step;
            const uint4 q = qs[u];
            const uint32_t n = q.z, kind = q.w & 0xffu, shift = (q.w >> 8) & 0xffu, mix_bits = (q.w >> 16) & 0xffu;
            const int32_t mix_res = (int32_t)q.w >> 24;
            if (kind == CH_PAIR_V || j >= n) continue;
            uint8_t *out = reinterpret_cast<uint8_t *>(((uint64_t)q.y << 32) | q.x) + (size_t)j * stride;
            if (kind == CH_ZERO) {
                for (uint32_t cc = 0; cc < zero_chans; cc++) store_sample<DEPTH>(out + cc * bps, 0);
                continue;
            }
            int32_t l = bu[lane * kTilePitch + pr];
            uint32_t low = 0;
            if (shift) {
                const uint32_t W = (kind == CH_MONO) ? shift : 2u * shift;
                if (sh_mode[u] == 1) {
                    low = __funnelshift_l(bswap32(w1[u]), bswap32(w0[u]), sh_bit[u]) >> (32u - W);
                } else {
                    const uint4 q2 = lds_meta(metas + pr * (uint32_t)sizeof(FinMeta) + 16u);
                    BitPeek bp;
                    bp.start(reinterpret_cast<const uint8_t *>(((uint64_t)q2.y << 32) | q2.x), q2.z);
                    low = bp.bits_at(q2.w + j * W, W);
                }
            }
            if (kind == CH_MONO) {
                if (shift) l = (int32_t)(((uint32_t)l << shift) | low);                     // :436-495
                store_sample<DEPTH>(out, l);
            } else {
                const int32_t v = bv[lane * kTilePitch + pr];
                int32_t r;
                if (mix_res != 0) {                         // :193-223
                    l = l + v - ((mix_res * v) >> (mix_bits & 31u));    // an out-of-range mixBits shifts like the reference's int32 shift on x86
                    r = l - v;
                } else {
                    r = v;
                }
                if (shift) {                                // :282-383
                    l = (int32_t)(((uint32_t)l << shift) | (low >> shift));
                    r = (int32_t)(((uint32_t)r << shift) | (low & ((1u << shift) - 1u)));
                }
                if (DEPTH == 16 && out_pair32) {
                    *reinterpret_cast<uint32_t *>(out) = ((uint32_t)l & 0xffffu) | ((uint32_t)r << 16);
                } else {
                    store_sample<DEPTH>(out, l);
                    store_sample<DEPTH>(out + bps, r);
                }
            }
        }
    }
}

// are all packets of the group regular?  (lanes past the last packet count as regular)
__device__ __forceinline__ bool group_is_regular(const DecArgs &A, uint32_t group, uint32_t lane)
{
    if (!A.fused) return false;
    const uint32_t slot = group * 32u + lane;
    const uint32_t cls = slot < A.num_packets ? A.pkt_class[A.perm[A.pkt_base + slot]] : 0u;
    return __all_sync(0xffffffffu, cls < kRegularClasses);
}

// ---- entropy kernel: one lane per packet ---------------------------------------------------------------------
// The serial walk through the packet's bits (codec/ALACDecoder.cu:571-1002): element loop, headers, Golomb
// streams, escape samples.  Residuals go to the channel tiles, headers to chan_hdr / chan_meta.
template <int DEPTH>
__global__ void __launch_bounds__(kRingStride) dec_entropy_kernel(DecArgs A)
{
    __shared__ uint32_t s_ring[kRingSlots][kRingStride];
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    static_assert(kRingStride == 32, "one group of 32 packets per CTA");
    if (group_is_regular(A, blockIdx.x, threadIdx.x)) return;          // dec_fused_kernel's group
    if (tid >= A.num_packets) return;
    const uint32_t pkt = A.perm[A.pkt_base + tid];
    const uint32_t nch = A.num_channels;
    const uint32_t size = A.pkt_size[pkt];
    const uint32_t cap_bits = size * 8u;
    const uint32_t slot_samples = A.pkt_samples[pkt];
    const uint32_t F = A.frame_length;
    int32_t *tile0 = A.chan_scratch + ((size_t)(tid >> 5) * nch * F) * 32u + (tid & 31u);
    DecChanMeta *meta = A.chan_meta + (size_t)pkt * nch;
    DecChanHdr *hdrs = A.chan_hdr + (size_t)pkt * nch;
    for (uint32_t c = 0; c < nch; c++) {            // channels that never arrive are zero-filled (:972-998)
        DecChanMeta z;
        z.n = slot_samples; z.shift_pos = 0; z.kind = CH_ZERO; z.shift = 0; z.mix_res = 0; z.mix_bits = 0;
        meta[c] = z;
    }
    BitPeek br;             // headers, escape samples
    br.start(A.packets + A.pkt_off[pkt], size);
    BitReader gr;           // Golomb streams
    gr.start(A.packets + A.pkt_off[pkt], size, &s_ring[0][threadIdx.x]);

    uint32_t n = F;
    uint32_t channel_index = 0;
    int32_t status = 0;
    ChanHeader hu, hv;

    while (status == 0) {
        if (!((br.pos >> 3) < size)) { status = -50; break; }                   // :615
        const uint32_t tag = br.get(3);
        if (tag == ID_SCE || tag == ID_LFE || tag == ID_CPE) {
            const bool pair = (tag == ID_CPE);
            if (pair && channel_index + 2 > nch) break;                         // :759-760
            if (!pair && channel_index >= nch) { status = -50; break; }
            br.pos += 4;                                                         // element instance tag
            if (br.get(12) != 0) { status = -50; break; }                       // :633
            const uint32_t hb = br.get(4);
            const uint32_t partial = hb >> 3;
            const uint32_t bytes_shifted = (hb >> 1) & 3u;
            if (bytes_shifted == 3) { status = -50; break; }                    // :641
            const uint32_t escape = hb & 1u;
            if (partial) { n = br.get(16) << 16; n |= br.get(16); }             // :650-654
            if (n > slot_samples) { status = -50; break; }
            DecChanMeta mu;
            mu.n = n; mu.shift_pos = 0; mu.kind = pair ? CH_PAIR_U : CH_MONO; mu.shift = 0; mu.mix_res = 0; mu.mix_bits = 0;
            int32_t *du = tile0 + (size_t)channel_index * F * 32u;
            int32_t *dv = du + (size_t)F * 32u;
            if (!escape) {
                const uint32_t chan_bits = DEPTH - bytes_shifted * 8 + (pair ? 1u : 0u);
                mu.mix_bits = (uint8_t)br.get(8);
                mu.mix_res = (int8_t)br.get(8);
                read_chan_header(br, hu);
                if (pair) read_chan_header(br, hv);
                if (bytes_shifted) {                                            // :675-679, :818-822
                    mu.shift = (uint8_t)(bytes_shifted * 8);
                    mu.shift_pos = br.pos;
                    br.pos += mu.shift * (pair ? 2u : 1u) * n;
                }
                store_chan_hdr(&hdrs[channel_index], hu, chan_bits);
                status = entropy_channel(gr, br, cap_bits, A, n, chan_bits, hu, du);
                if (status == 0 && pair) {
                    store_chan_hdr(&hdrs[channel_index + 1], hv, chan_bits);
                    status = entropy_channel(gr, br, cap_bits, A, n, chan_bits, hv, dv);
                }
            } else {
                // uncompressed element (:697-727, :856-896): raw samples, pairs interleaved
                const uint32_t sh = 32u - DEPTH;
                hdrs[channel_index].num = (uint8_t)kRawChannel;
                if (pair) hdrs[channel_index + 1].num = (uint8_t)kRawChannel;
                for (uint32_t j = 0; j < n; j++) {
                    du[(size_t)j * 32u] = (int32_t)(br.get(DEPTH) << sh) >> sh;
                    if (pair) dv[(size_t)j * 32u] = (int32_t)(br.get(DEPTH) << sh) >> sh;
                }
            }
            if (status) break;
            meta[channel_index] = mu;
            if (pair) { mu.kind = CH_PAIR_V; meta[channel_index + 1] = mu; }
            channel_index += pair ? 2u : 1u;
        } else if (tag == ID_CCE || tag == ID_PCE) {
            status = -50;                                                       // :932-939
        } else if (tag == ID_DSE) {                                             // :1033-1059
            br.pos += 4;
            const uint32_t align = br.get(1);
            uint32_t count = br.get(8);
            if (count == 255) count += br.get(8);
            if (align && (br.pos & 7u)) br.pos += 8u - (br.pos & 7u);
            br.pos += count * 8;
            if ((br.pos >> 3) > size) status = -50;
        } else if (tag == ID_FIL) {                                             // :1012-1027
            int32_t count = (int32_t)br.get(4);
            if (count == 15) count += (int32_t)br.get(8) - 1;
            br.pos += (uint32_t)count * 8;
            if ((br.pos >> 3) > size) status = -50;
        } else {
            break;                                                              // ID_END :955-961
        }
        if (channel_index >= nch) break;                                        // :966-967
    }
    cp_async_wait<0>();
    A.pkt_status[pkt] = status;
}

// ---- finish kernel: predictor + un-mix + output, one lane per (packet, element) -------------------------------------
// unpc_block (codec/dp_dec.c:55-381), then unmixNN / copyPredictorToNN (codec/ALACDecoder.cu:193-495).
// A CTA is one (group of 32 packets, channel slot c) and has two warps, lane = packet: warp 0 predicts slot c (a mono
// channel or the U of a pair), warp 1 the V of the pair in slot c + 1.  CTAs of slots that are a V in every packet
// exit at once (their U's CTA does the work).
// The frame is walked in tiles of 32 samples:
//   1. the tile's residual rows (128 B each: 32 lanes) arrive in shared memory by 16-byte cp.async, one tile
//      ahead of the arithmetic (a whole tile of predictor work hides the HBM latency);
//   2. serial phase: every lane runs the predictor down its own column, in place (conflict-free: bank == lane);
//   3. parallel phase: the roles flip -- lane = sample, each warp loops over half of the 32 packets -- so a
//      packet's 32 finished sample-frames are un-mixed, merged with their shift bytes, packed and stored as one
//      contiguous run.
// The class permutation makes the lanes of a warp run the same tap counts.
enum : uint32_t { PM_PASS = 0, PM_FAST4 = 1, PM_FAST8 = 2, PM_WRAP = 4 };
struct PredState { int32_t a[8]; int32_t hist[9]; };

// rows [r0, r1) of the lane's tile column; absolute sample index of row r is j0 + r
template <int TAPS, bool WRAP>
__device__ __forceinline__ void unpc_rows(PredState &s, int32_t *col, uint32_t j0, uint32_t r0, uint32_t r1, uint32_t chanshift)
{
    int32_t a[TAPS], hist[TAPS + 1];
#pragma unroll
    for (int k = 0; k < TAPS; k++) a[k] = s.a[k];
#pragma unroll
    for (int k = 0; k <= TAPS; k++) hist[k] = s.hist[k];
    uint32_t r = r0;
    // warm-up samples 0..TAPS: first differences (codec/dp_dec.c:65, :97-101)
    for (; r < r1 && j0 + r <= (uint32_t)TAPS; r++) {
        const int32_t res = col[r * kTilePitch];
        const int32_t x = (j0 + r) ? sext_bits(res + hist[0], chanshift) : res;
        col[r * kTilePitch] = x;
#pragma unroll
        for (int k = TAPS; k > 0; k--) hist[k] = hist[k - 1];
        hist[0] = x;
    }
#pragma unroll 2
    for (; r < r1; r++) col[r * kTilePitch] = predict_dec_step<TAPS, WRAP>(col[r * kTilePitch], hist, a, chanshift);
#pragma unroll
    for (int k = 0; k < TAPS; k++) s.a[k] = a[k];
#pragma unroll
    for (int k = 0; k <= TAPS; k++) s.hist[k] = hist[k];
}

__device__ __forceinline__ void unpc_rows_any(uint32_t mode, PredState &s, int32_t *col, uint32_t j0, uint32_t r0, uint32_t r1, uint32_t chanshift)
{
    switch (mode) {
    case PM_FAST4: unpc_rows<4, false>(s, col, j0, r0, r1, chanshift); break;
    case PM_FAST8: unpc_rows<8, false>(s, col, j0, r0, r1, chanshift); break;
    case PM_FAST4 | PM_WRAP: unpc_rows<4, true>(s, col, j0, r0, r1, chanshift); break;
    case PM_FAST8 | PM_WRAP: unpc_rows<8, true>(s, col, j0, r0, r1, chanshift); break;
    default: break;
    }
}

// any numactive 0..31, any denShift, mode != 0 (codec/dp_dec.c:67-95, :335-380; codec/ALACDecoder.cu:686-694):
// in place on the lane's column of the channel tile in global memory
static __device__ __noinline__ void unpc_general(int32_t *col, uint32_t n, DecChanHdr h, uint32_t chanshift)
{
    int32_t ring[32];
    for (int k = 0; k < 32; k++) ring[k] = 0;
    const int32_t num = (int32_t)h.num;
    const uint32_t ds = h.den_shift;
    const int32_t half = ds ? (1 << (ds - 1)) : 0;
    int32_t pre = 0;        // running value of the mode != 0 first-difference pass
    int32_t prev = 0;
    for (uint32_t j = 0; j < n; j++) {
        int32_t r = col[(size_t)j * 32u];
        if (h.mode != 0) {                          // unpc_block(pred, pred, n, nil, 31, chanBits, 0)
            r = j ? sext_bits(r + pre, chanshift) : r;
            pre = r;
        }
        int32_t x;
        if (j == 0 || num == 0) {
            x = r;
        } else if (num == 31 || j <= (uint32_t)num) {
            x = sext_bits(r + prev, chanshift);
        } else {
            const int32_t top = ring[(j - num - 1) & 31u];
            int32_t acc = 0;
            for (int32_t k = 0; k < num; k++) acc += (int32_t)h.coefs[k] * (ring[(j - 1 - k) & 31u] - top);
            x = sext_bits(r + top + ((acc + half) >> ds), chanshift);
            int32_t left = r;
            if (r > 0) {
                for (int32_t k = num - 1; k >= 0; k--) {
                    const int32_t dd = top - ring[(j - 1 - k) & 31u];
                    const int32_t s = sign3(dd);
                    h.coefs[k] = (int16_t)(h.coefs[k] - s);
                    left -= (num - k) * ((s * dd) >> ds);
                    if (left <= 0) break;
                }
            } else if (r < 0) {
                for (int32_t k = num - 1; k >= 0; k--) {
                    const int32_t dd = top - ring[(j - 1 - k) & 31u];
                    const int32_t s = sign3(dd);
                    h.coefs[k] = (int16_t)(h.coefs[k] + s);
                    left -= (num - k) * ((-s * dd) >> ds);
                    if (left >= 0) break;
                }
            }
        }
        ring[j & 31u] = x;
        prev = x;
        col[(size_t)j * 32u] = x;
    }
}

// per-lane predictor set-up for one channel; returns the PM_* mode.  Channels the register paths do not cover
// are finished right here, in place in global memory, and then pass through the tiles untouched.
template <bool ALLOW_GENERAL = true>
__device__ __forceinline__ uint32_t pred_setup(const DecChanHdr *hp, uint32_t n, int32_t *gcol, PredState &s)
{
    const uint32_t num = hp->num;
#pragma unroll
    for (int k = 0; k < 9; k++) s.hist[k] = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s.a[k] = 0;
    if (num == kRawChannel) return PM_PASS;
    const uint32_t chanshift = 32u - hp->chan_bits;
    if (hp->mode == 0 && hp->den_shift == kDenShift && (num == 4 || num == 8)) {
        // coefficients move by at most 1 per sample: if max|a| + n stays inside int16 no update can wrap and
        // the cheaper no-wrap step is exact
        int32_t amax = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if ((uint32_t)k < num) { s.a[k] = hp->coefs[k]; amax = max(amax, abs(s.a[k])); }
        }
        const bool safe = (uint32_t)amax + n <= 32767u;
        return (num == 4 ? PM_FAST4 : PM_FAST8) | (safe ? 0u : (uint32_t)PM_WRAP);
    }
    if (ALLOW_GENERAL) {
        unpc_general(gcol, n, *hp, chanshift);
        __threadfence();    // other lanes of the warp copy these words into the tiles
    }
    return PM_PASS;
}

template <int DEPTH>
__global__ void __launch_bounds__(64) dec_finish_kernel(DecArgs A)
{
    __shared__ __align__(16) int32_t s_tile[2][2][kTileWords];     // [buffer][channel c / c + 1][row * pitch + lane]
    __shared__ FinMeta s_meta[32];
        const uint32_t nch = A.num_channels, F = A.frame_length;
    const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31u;      // warp w owns channel slot c + w
    const uint32_t group = blockIdx.x / nch, c = blockIdx.x - group * nch;
    const uint32_t slot = group * 32u + lane;
    // 16-bit stereo: one 32-bit store per sample-frame when the output is 4-byte aligned
    const bool out_pair32 = (nch == 2) && ((reinterpret_cast<uintptr_t>(A.pcm_out) & 3u) == 0);
    if (group_is_regular(A, group, lane)) return;                       // dec_fused_kernel's group

    // ---- this lane's packet: what is channel slot c?  (both warps read the same records)
    constexpr uint32_t bps = DepthTraits<DEPTH>::kBytes;
    FinMeta M;
    M.kind = CH_PAIR_V; M.n = 0; M.out0 = nullptr; M.pkt = nullptr; M.pkt_size = 0; M.shift_pos = 0; M.shift = 0; M.mix_bits = 0; M.mix_res = 0;
    uint32_t pkt = 0;
    if (slot < A.num_packets) {
        pkt = A.perm[A.pkt_base + slot];
        const DecChanMeta m = A.chan_meta[(size_t)pkt * nch + c];
        M.kind = m.kind; M.n = m.n; M.shift_pos = m.shift_pos; M.shift = m.shift; M.mix_bits = m.mix_bits; M.mix_res = m.mix_res;
        M.out0 = A.pcm_out + A.out_frame[pkt] * (nch * bps) + (size_t)c * bps;
        M.pkt = A.packets + A.pkt_off[pkt];
        M.pkt_size = A.pkt_size[pkt];
    }
    if (!__any_sync(0xffffffffu, M.kind != CH_PAIR_V)) return;     // a slot that is a V everywhere: done by its U's CTA (both warps agree)
    if (w == 0) s_meta[lane] = M;

    // which channel does this lane predict?  warp 0: slot c (mono or U), warp 1: slot c + 1 (V of a pair)
    const bool mine = (w == 0) ? (M.kind == CH_MONO || M.kind == CH_PAIR_U) : (M.kind == CH_PAIR_U);
    int32_t *gtile = A.chan_scratch + ((size_t)(group * nch + c + w) * F) * 32u;
    PredState st;
    uint32_t mode = PM_PASS, chanshift = 0;
    if (mine) {
        const DecChanHdr *hp = A.chan_hdr + (size_t)pkt * nch + c + w;
        chanshift = 32u - hp->chan_bits;
        mode = pred_setup(hp, M.n, gtile + lane, st);
    }
    const uint32_t metas_saddr = (uint32_t)__cvta_generic_to_shared(s_meta);
    const bool warp_has_tiles = __any_sync(0xffffffffu, mine || (w == 0 && M.kind != CH_PAIR_V));
    const uint32_t n_pred = mine ? M.n : 0u;
    const uint32_t n_max = __reduce_max_sync(0xffffffffu, (M.kind == CH_PAIR_V) ? 0u : M.n);      // same in both warps
    const uint32_t tiles = (n_max + kTileRows - 1) / kTileRows;
    __syncthreads();

    // a tile = 32 rows x 8 chunks of 16 bytes; lane i takes chunks i, i + 32, ...
    auto request = [&](uint32_t t) {
        int32_t *buf = s_tile[t & 1u][w];
        if (warp_has_tiles) {
#pragma unroll
            for (uint32_t i = 0; i < 8; i++) {
                const uint32_t ch = i * 32u + lane, row = ch >> 3, part = ch & 7u;
                const uint32_t j = t * kTileRows + row;
                const bool in = j < F && t < tiles;
                const size_t goff = (size_t)(in ? j : 0u) * 32u + part * 4u;
                cp_async_16((uint32_t)__cvta_generic_to_shared(buf + row * kTilePitch + part * 4u), gtile + goff, in ? 16u : 0u);
            }
        }
        cp_async_commit();
    };
    request(0);
    request(1);
    for (uint32_t t = 0; t < tiles; t++) {
        cp_async_wait<1>();             // tile t is in (only tile t + 1 may still be in flight)
        __syncwarp();
        int32_t *bu = s_tile[t & 1u][0], *bv = s_tile[t & 1u][1];
        const uint32_t j0 = t * kTileRows;
        // ---- serial phase: lane = packet, each warp on its own channel
        if (j0 < n_pred) unpc_rows_any(mode, st, s_tile[t & 1u][w] + lane, j0, 0, min(n_pred - j0, kTileRows), chanshift);
        __syncthreads();
        // ---- parallel phase: lane = sample; warp w takes the packets of its parity
        flush_tile<DEPTH>(A, metas_saddr, bu, bv, j0, lane, w, 2, out_pair32, 1, false);
        __syncthreads();
        request(t + 2);                 // refills the buffer just drained (an empty group past the last tile)
    }
    cp_async_wait<0>();
}

// ---- fused decode kernel: regular mono / stereo packets ---------------------------------------------------------------
// A CTA is one group of 32 packets and two warps that run concurrently:
//   warp 0 (entropy): lane = packet.  Parses the element header, then decodes the Golomb streams of channel 0 and
//           channel 1 tile by tile (32 samples x 32 packets) into shared memory;
//   warp 1 (finish):  consumes each tile as it becomes ready: predictor down each lane's column (lane = packet), then,
//           with the roles flipped (lane = sample), un-mix / shift-merge / pack / store as in dec_finish_kernel.
// The two serial chains of a packet (entropy coder and predictor) thus overlap instead of adding up, and the residuals
// never leave the SM.  Only the finished U samples of a stereo pair make one trip through the channel scratch, because
// the whole U stream precedes the V stream in the packet.  Hand-off uses two residual buffers and four named barriers
// (FULL / EMPTY per buffer, 64 participants each: 32 arrive, 32 wait).
// "Regular" (classes 0..3 of dec_header_kernel) means: exactly one element that matches the channel count, compressed,
// predictor mode 0, denShift 9, 4 or 8 taps.  Every lane of the entropy warp then executes the same tile loop, which the
// barrier protocol needs; groups holding any other packet are left to dec_entropy_kernel + dec_finish_kernel.
template <int DEPTH>
__global__ void __launch_bounds__(64) dec_fused_kernel(DecArgs A)
{
    __shared__ uint32_t s_ring[kRingSlots][32];
    __shared__ __align__(16) int32_t s_res[2][kTileWords];      // residual tiles: entropy warp -> finish warp
    __shared__ __align__(16) int32_t s_xu[kTileWords];          // finished U tile, back from the scratch (stereo, V phase)
    __shared__ FinMeta s_meta[32];
    __shared__ uint32_t s_nmax;

    const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t group = blockIdx.x;
    if (!group_is_regular(A, group, lane)) return;              // both warps agree
    const uint32_t nch = A.num_channels, F = A.frame_length;    // 1 or 2 here
    const uint32_t slot = group * 32u + lane;
    const bool valid = slot < A.num_packets;
    const uint32_t pkt = valid ? A.perm[A.pkt_base + slot] : 0u;
    int32_t *gtile_u = A.chan_scratch + ((size_t)(group * nch) * F) * 32u;
    const bool out_pair32 = (nch == 2) && ((reinterpret_cast<uintptr_t>(A.pcm_out) & 3u) == 0);

    if (w == 0) {
        // ================= entropy warp =================
        const uint32_t size = valid ? A.pkt_size[pkt] : 0u;
        const uint32_t cap_bits = size * 8u;
        const uint32_t slot_samples = valid ? A.pkt_samples[pkt] : 0u;
        const uint8_t *packet = A.packets + (valid ? A.pkt_off[pkt] : 0u);
        int32_t status = 0;
        uint32_t n = 0;
        ChanHeader hu, hv;
        hu.pb_factor = hv.pb_factor = 4;
        uint32_t chan_bits = DEPTH;
        FinMeta M;
        M.kind = valid ? (uint8_t)CH_ZERO : (uint8_t)CH_PAIR_V; M.n = slot_samples;
        M.out0 = A.pcm_out + (valid ? A.out_frame[pkt] : 0u) * (nch * DepthTraits<DEPTH>::kBytes);
        M.pkt = packet; M.pkt_size = size; M.shift_pos = 0; M.shift = 0; M.mix_bits = 0; M.mix_res = 0;
        BitPeek bp;
        bp.start(packet, size);
        if (valid) {
            // the same walk as dec_entropy_kernel up to the first audio element (codec/ALACDecoder.cu:571-694)
            for (int guard = 0; guard < 64 && status == 0; guard++) {
                if (!((bp.pos >> 3) < size)) { status = -50; break; }
                const uint32_t tag = bp.get(3);
                if (tag == ID_SCE || tag == ID_LFE || tag == ID_CPE) {
                    const bool pair = (tag == ID_CPE);
                    bp.pos += 4;
                    if (bp.get(12) != 0) { status = -50; break; }
                    const uint32_t hb = bp.get(4);
                    const uint32_t bytes_shifted = (hb >> 1) & 3u;
                    n = F;
                    if (hb >> 3) { n = bp.get(16) << 16; n |= bp.get(16); }
                    if (bytes_shifted == 3 || (hb & 1u) || n > slot_samples || pair != (nch == 2)) { status = -50; break; }   // excluded by the class
                    chan_bits = DEPTH - bytes_shifted * 8 + (pair ? 1u : 0u);
                    M.mix_bits = (uint8_t)bp.get(8);
                    M.mix_res = (int8_t)bp.get(8);
                    read_chan_header(bp, hu);
                    if (pair) read_chan_header(bp, hv);
                    if (bytes_shifted) {
                        M.shift = (uint8_t)(bytes_shifted * 8);
                        M.shift_pos = bp.pos;
                        bp.pos += M.shift * (pair ? 2u : 1u) * n;
                    }
                    M.kind = pair ? CH_PAIR_U : CH_MONO;
                    M.n = n;
                    break;
                } else if (tag == ID_DSE) {
                    bp.pos += 4;
                    const uint32_t align = bp.get(1);
                    uint32_t count = bp.get(8);
                    if (count == 255) count += bp.get(8);
                    if (align && (bp.pos & 7u)) bp.pos += 8u - (bp.pos & 7u);
                    bp.pos += count * 8;
                    if ((bp.pos >> 3) > size) status = -50;
                } else if (tag == ID_FIL) {
                    int32_t count = (int32_t)bp.get(4);
                    if (count == 15) count += (int32_t)bp.get(8) - 1;
                    bp.pos += (uint32_t)count * 8;
                    if ((bp.pos >> 3) > size) status = -50;
                } else {
                    status = -50;
                }
            }
            if (status) { n = 0; M.kind = CH_ZERO; M.n = slot_samples; }      // zero-filled like any failed packet
            else {
                DecChanHdr *hdrs = A.chan_hdr + (size_t)pkt * nch;
                store_chan_hdr(&hdrs[0], hu, chan_bits);
                if (nch == 2) store_chan_hdr(&hdrs[1], hv, chan_bits);
            }
        }
        s_meta[lane] = M;
        const uint32_t n_max = __reduce_max_sync(0xffffffffu, valid ? M.n : 0u);     // failed packets are zero-filled to their slot size
        if (lane == 0) s_nmax = n_max;
        __threadfence_block();
        __syncthreads();                                        // (1) headers and metas are out
        const uint32_t tiles = (n_max + kTileRows - 1) / kTileRows;
        BitReader br;
        br.start(packet, size, &s_ring[0][lane]);
        br.seek(bp.pos);
        for (uint32_t c = 0; c < nch; c++) {
            AgDec ag;
            ag.start(br, n, A.mb, (A.pb * (c ? hv.pb_factor : hu.pb_factor)) / 4, A.kb, chan_bits);     // codec/ALACDecoder.cu:682
            for (uint32_t t = 0; t < tiles; t++) {
                const uint32_t b = t & 1u;
                __syncwarp();
                named_sync<BAR_EMPTY0>(b != 0);
                int32_t *col = s_res[b] + lane;
                const uint32_t j0 = t * kTileRows;
                // (not unrolled further: at() is long and the kernel is sensitive to instruction-cache misses)
#pragma unroll 1
                for (uint32_t r0 = 0; r0 < kTileRows; r0 += kTopUpEvery) {
#pragma unroll 1
                    for (uint32_t i = 0; i < kTopUpEvery; i++) col[(r0 + i) * kTilePitch] = ag.at(br, j0 + r0 + i);
                    br.top_up();
                }
                __syncwarp();
                named_arrive<BAR_FULL0>(b != 0);
            }
            ag.finish(br, cap_bits);
            if (n && !status) status = ag.status;
            __syncthreads();                                    // (2) the finish warp is done with this channel's tiles
        }
        cp_async_wait<0>();
        if (valid) A.pkt_status[pkt] = status;
        return;
    }

    // ================= finish warp =================
    __syncthreads();                                            // (1)
    const uint32_t n_max = s_nmax;
    const uint32_t tiles = (n_max + kTileRows - 1) / kTileRows;
    const FinMeta M = s_meta[lane];
    const uint32_t metas_saddr = (uint32_t)__cvta_generic_to_shared(s_meta);
    // every packet of the group a stereo pair without shift bytes (slots past the last packet: n = 0), aligned 16-bit output
    const bool simple16 = DEPTH == 16 && out_pair32 &&
                          __all_sync(0xffffffffu, (M.kind == CH_PAIR_U && M.shift == 0) || (M.kind == CH_PAIR_V && M.n == 0));
    const bool mine = (M.kind == CH_MONO || M.kind == CH_PAIR_U);
    const uint32_t n_pred = mine ? M.n : 0u;
    for (uint32_t c = 0; c < nch; c++) {
        PredState st;
        uint32_t mode = PM_PASS, chanshift = 0;
        if (mine) {
            const DecChanHdr *hp = A.chan_hdr + (size_t)pkt * nch + c;
            chanshift = 32u - hp->chan_bits;
            mode = pred_setup<false>(hp, M.n, nullptr, st);     // regular: a register path, never the general one
        }
        const bool last_chan = (c + 1 == nch);
        for (uint32_t b = 0; b < min(tiles, 2u); b++) named_arrive<BAR_EMPTY0>(b != 0);   // both buffers start empty
        for (uint32_t t = 0; t < tiles; t++) {
            const uint32_t b = t & 1u;
            const uint32_t j0 = t * kTileRows;
            if (nch == 2 && last_chan) {
                // the finished U tile comes back from the scratch while this warp waits for the V residuals
#pragma unroll
                for (uint32_t i = 0; i < 8; i++) {
                    const uint32_t ch = i * 32u + lane, row = ch >> 3, part = ch & 7u;
                    const bool in = j0 + row < F;
                    cp_async_16((uint32_t)__cvta_generic_to_shared(s_xu + row * kTilePitch + part * 4u),
                                gtile_u + (size_t)(in ? j0 + row : 0u) * 32u + part * 4u, in ? 16u : 0u);
                }
                cp_async_commit();
            }
            __syncwarp();
            named_sync<BAR_FULL0>(b != 0);
            int32_t *buf = s_res[b];
            if (j0 < n_pred) unpc_rows_any(mode, st, buf + lane, j0, 0, min(n_pred - j0, kTileRows), chanshift);
            if (!last_chan) {
                // U of a pair: park the finished column in the scratch (one 128-byte row per store)
                int32_t *g = gtile_u + (size_t)j0 * 32u + lane;
                if (j0 + kTileRows <= F) {
#pragma unroll
                    for (uint32_t r = 0; r < kTileRows; r++) g[r * 32u] = buf[r * kTilePitch + lane];
                } else {
                    for (uint32_t r = 0; r < kTileRows && j0 + r < F; r++) g[(size_t)r * 32u] = buf[r * kTilePitch + lane];
                }
            } else {
                cp_async_wait<0>();
                __syncwarp();
                flush_tile<DEPTH>(A, metas_saddr, nch == 2 ? s_xu : buf, buf, j0, lane, 0, 1, out_pair32, nch, simple16);
                __syncwarp();
            }
            __syncwarp();
            if (t + 2 < tiles) named_arrive<BAR_EMPTY0>(b != 0);
        }
        __threadfence_block();
        __syncthreads();                                        // (2)
    }
}

// ---- CAF packet table on the device -------------------------------------------------------
// The 'pakt' chunk stores each packet size as a BER integer (7 bits per byte, high bit = "more",
// convert-utility/CAFFileALAC.cpp:189-258).  ber_flag_kernel marks the last byte of every entry; an exclusive
// scan of the marks numbers the entries; ber_value_kernel assembles each entry from the <= 5 bytes that end at
// its mark; ber_count_kernel finds where the reference's decode loop would stop (a zero size, or a packet that
// no longer fits the data chunk, convert-utility/main.cu:717).
static __global__ void ber_flag_kernel(const uint8_t *table, uint64_t nbytes, uint32_t *flags)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nbytes) flags[i] = (table[i] & 0x80u) ? 0u : 1u;
}

static __global__ void ber_value_kernel(const uint8_t *table, uint64_t nbytes, const uint64_t *entry_of_byte, uint32_t *sizes, uint64_t cap)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbytes || (table[i] & 0x80u)) return;
    const uint64_t k = entry_of_byte[i];
    if (k >= cap) return;
    uint32_t v = table[i] & 0x7fu, shift = 7, len = 1;
    for (uint64_t j = i; j > 0 && (table[j - 1] & 0x80u); j--) {
        if (++len > 5) { v = 0; break; }                    // ReadBERInteger gives up after 5 bytes (:246-250)
        v |= (uint32_t)(table[j - 1] & 0x7fu) << shift;
        shift += 7;
    }
    sizes[k] = v;
}

static __global__ void ber_count_kernel(const uint32_t *sizes, const uint64_t *offsets, uint64_t entries, uint64_t data_bytes,
                                 unsigned long long *first_bad)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= entries) return;
    if (sizes[k] == 0 || offsets[k] + sizes[k] > data_bytes) atomicMin(first_bad, (unsigned long long)k);
}

// the other direction: packet sizes -> BER bytes (convert-utility/CAFFileALAC.cpp:189-222 WriteBERInteger).
// ber_len_kernel gives each entry's byte count (1..5), an exclusive scan places it, ber_emit_kernel writes it.
static __global__ void ber_len_kernel(const uint32_t *sizes, uint64_t n, uint32_t *lens)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t v = sizes[k];
    lens[k] = v < (1u << 7) ? 1u : v < (1u << 14) ? 2u : v < (1u << 21) ? 3u : v < (1u << 28) ? 4u : 5u;
}

static __global__ void ber_emit_kernel(const uint32_t *sizes, const uint64_t *offsets, uint64_t n, uint8_t *table)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t v = sizes[k];
    const uint32_t len = (uint32_t)(offsets[k + 1] - offsets[k]);
    uint8_t *p = table + offsets[k];
    for (uint32_t i = 0; i < len; i++) {
        const uint32_t sh = 7u * (len - 1u - i);
        p[i] = (uint8_t)(((v >> sh) & 0x7fu) | (i + 1u < len ? 0x80u : 0u));
    }
}


// ---- host-side launchers (one translation unit per bit depth, see alac_encode.cuh) ---------------------------------
// The decode kernels keep their tiles / rings in static shared memory and want 8-9 CTAs per SM resident (one
// wave for the 1-hour workload).  With the default carve-out the driver leaves most of the 228 KB to L1 and only
// about half of those CTAs fit, so ask for the maximum shared-memory carve-out once per device.
template <int DEPTH>
void dec_configure()
{
    cudaFuncSetAttribute(dec_fused_kernel<DEPTH>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(dec_finish_kernel<DEPTH>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(dec_entropy_kernel<DEPTH>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

// the three main kernels of one chunk; ev: optional {after fused, after entropy} events.  Returns kernels launched.
template <int DEPTH>
uint32_t dec_launch_main(cudaStream_t s, const DecArgs &A, cudaEvent_t *ev)
{
    const uint32_t lgrid = (A.num_packets + kRingStride - 1) / kRingStride;
    const uint32_t groups = (A.num_packets + 31) / 32;
    uint32_t launches = 2;
    // regular mono / stereo groups: entropy and finish warps side by side in one kernel
    if (A.fused) { dec_fused_kernel<DEPTH><<<groups, 64, 0, s>>>(A); launches++; }
    if (ev) cudaEventRecord(ev[0], s);
    // everything else (multichannel, escapes, other predictor set-ups): the two general kernels; groups the
    // fused kernel took return at once
    dec_entropy_kernel<DEPTH><<<lgrid, kRingStride, 0, s>>>(A);
    if (ev) cudaEventRecord(ev[1], s);
    dec_finish_kernel<DEPTH><<<groups * A.num_channels, 64, 0, s>>>(A);
    return launches;
}

#ifndef ALAC_INSTANTIATE_DEPTH
#define ALAC_DEC_EXTERN(D)                                                              \
    extern template void dec_configure<D>();                                             \
    extern template uint32_t dec_launch_main<D>(cudaStream_t, const DecArgs &, cudaEvent_t *);
ALAC_DEC_EXTERN(16) ALAC_DEC_EXTERN(20) ALAC_DEC_EXTERN(24) ALAC_DEC_EXTERN(32)
#undef ALAC_DEC_EXTERN
#endif

}  // namespace alacb
