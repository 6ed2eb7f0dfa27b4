// alac_decode.cuh -- decode kernels.
//
//   dec_header_kernel   one lane per packet: reads the first audio element's header to learn the
//                       packet's sample count (partial-frame field), so output offsets can be scanned.
//                       Also classifies the packet by predictor orders; dec_perm_kernel turns the
//                       classes into a lane -> packet permutation so warps run uniform tap counts.
//   dec_lane_kernel     one lane per packet: the serial part of ALACDecoder::Decode
//                       (codec/ALACDecoder.cu:571-1002): element loop, header parse, dyn_decomp and
//                       unpc_block fused per sample.  Each channel's int32 samples (u / v / mono) go
//                       to a scratch laid out [group of 32 packets][channel][sample][lane], so every
//                       store of a warp is one 128-byte line; a DecChanMeta per channel says how to
//                       finish it.  The bitstream arrives through a cp.async shared-memory ring.
//   dec_output_kernel   the data-parallel part (codec/ALACDecoder.cu:193-495 unmixNN /
//                       copyPredictorToNN): 32 packets x 32 samples tiles are transposed through
//                       shared memory, un-mixed, merged with the shift bytes read straight from the
//                       packet, packed and stored with coalesced writes.
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
};

// how dec_output_kernel finishes one channel of one packet
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

// walk element tags until the first SCE/LFE/CPE and return its sample count
__global__ void dec_header_kernel(DecArgs A)
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
            br.pos += 4 + 12;
            const uint32_t hb = br.get(4);
            if (hb >> 3) { n = br.get(16) << 16; n |= br.get(16); }
            if (!(hb & 1u)) {
                // predictor orders of the first element: class = 3 * order(U) + order(V), order in {4, 8, other}
                br.pos += 16 + 8;
                const uint32_t nu = br.get(8) & 0x1fu;
                uint32_t nv = 4;
                if (tag == ID_CPE) { br.pos += nu * 16 + 8; nv = br.get(8) & 0x1fu; }
                cls = 3 * (nu == 4 ? 0u : nu == 8 ? 1u : 2u) + (nv == 4 ? 0u : nv == 8 ? 1u : 2u);
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
__global__ void dec_perm_kernel(DecArgs A)
{
    const uint32_t local = blockIdx.x * blockDim.x + threadIdx.x;
    if (local >= A.num_packets) return;
    const uint32_t p = A.pkt_base + local;
    const uint32_t cls = A.pkt_class[p];
    uint32_t first = 0;
    for (uint32_t c = 0; c < cls; c++) first += A.class_count[c];
    A.perm[A.pkt_base + first + A.pkt_rank[p]] = p;
}

struct ChanHeader {
    uint32_t mode, den_shift, pb_factor, num;
    int16_t coefs[32];
};

__device__ __forceinline__ void read_chan_header(BitReader &br, ChanHeader &h)
{
    uint32_t hb = br.get(8);                        // codec/ALACDecoder.cu:660-669
    h.mode = hb >> 4;
    h.den_shift = hb & 0xfu;
    hb = br.get(8);
    h.pb_factor = hb >> 5;
    h.num = hb & 0x1fu;
    for (uint32_t i = 0; i < h.num; i++) h.coefs[i] = (int16_t)br.get(16);
}

// dyn_decomp + unpc_block for one channel, streamed; out(j, sample) receives the n samples.
template <int TAPS, bool WRAP, class Out>
__device__ __forceinline__ void decode_channel_fast(BitReader &br, uint32_t cap_bits, AgDec &ag, uint32_t n,
                                                    const ChanHeader &h, uint32_t chanshift, Out &out)
{
    int32_t a[TAPS], hist[TAPS + 1];
#pragma unroll
    for (int k = 0; k < TAPS; k++) a[k] = h.coefs[k];
#pragma unroll
    for (int k = 0; k <= TAPS; k++) hist[k] = 0;
    int32_t prev = 0;
    const uint32_t warm = min(n, (uint32_t)TAPS + 1u);
    for (uint32_t j = 0; j < warm; j++) {           // codec/dp_dec.c:65, :97-101
        const int32_t r = ag.next(br, cap_bits);
        const int32_t x = j ? sext_bits(r + prev, chanshift) : r;
        out(j, x);
#pragma unroll
        for (int k = TAPS; k > 0; k--) hist[k] = hist[k - 1];
        hist[0] = x;
        prev = x;
    }
    for (uint32_t j = TAPS + 1; j < n; j++) {
        const int32_t r = ag.next(br, cap_bits);
        out(j, predict_dec_step<TAPS, WRAP>(r, hist, a, chanshift));
    }
}

// any numactive 0..31, any denShift, mode != 0 (codec/dp_dec.c:67-95, :335-380; codec/ALACDecoder.cu:686-694)
// Everything is taken BY VALUE (and the reader / coder state handed back) so that this rarely-run,
// out-of-line routine does not force the hot objects of the fast paths into local memory.
struct GeneralState { BitReader br; AgDec ag; };
template <class Out>
__device__ __noinline__ GeneralState decode_channel_general(BitReader br, uint32_t cap_bits, AgDec ag, uint32_t n,
                                                            ChanHeader h, uint32_t chanshift, Out out)
{
    int32_t ring[32];
    for (int k = 0; k < 32; k++) ring[k] = 0;
    const int32_t num = (int32_t)h.num;
    const uint32_t ds = h.den_shift;
    const int32_t half = ds ? (1 << (ds - 1)) : 0;
    int32_t pre = 0;        // running value of the mode != 0 first-difference pass
    int32_t prev = 0;
    for (uint32_t j = 0; j < n; j++) {
        int32_t r = ag.next(br, cap_bits);
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
        out(j, x);
    }
    GeneralState gs;
    gs.br = br;
    gs.ag = ag;
    return gs;
}

template <class Out>
__device__ __forceinline__ int32_t decode_channel(BitReader &br, uint32_t cap_bits, const DecArgs &A, uint32_t n,
                                                  uint32_t chan_bits, ChanHeader &h, Out &out)
{
    AgDec ag;
    ag.start(br, n, A.mb, (A.pb * h.pb_factor) / 4, A.kb, chan_bits);       // codec/ALACDecoder.cu:682
    const uint32_t chanshift = 32u - chan_bits;
    // coefficients move by at most 1 per sample: if max|a| + n stays inside int16 no update can wrap and
    // the cheaper no-wrap step is exact
    int32_t amax = 0;
    for (uint32_t i = 0; i < h.num; i++) amax = max(amax, abs((int32_t)h.coefs[i]));
    const bool safe = (uint32_t)amax + n <= 32767u;
    const bool fast = h.mode == 0 && h.den_shift == kDenShift;
    if (fast && h.num == 4) {
        if (safe) decode_channel_fast<4, false>(br, cap_bits, ag, n, h, chanshift, out);
        else decode_channel_fast<4, true>(br, cap_bits, ag, n, h, chanshift, out);
    } else if (fast && h.num == 8) {
        if (safe) decode_channel_fast<8, false>(br, cap_bits, ag, n, h, chanshift, out);
        else decode_channel_fast<8, true>(br, cap_bits, ag, n, h, chanshift, out);
    } else {
        const GeneralState gs = decode_channel_general(br, cap_bits, ag, n, h, chanshift, out);
        br = gs.br;
        ag = gs.ag;
    }
    // dyn_decomp's exit check "cur <= end" (codec/ag_dec.c:359)
    if (!ag.status && (br.pos >> 3) > (cap_bits >> 3)) ag.status = -50;
    return ag.status;
}

// channel samples -> scratch [sample][lane]: one coalesced line per warp store
struct ScratchOut {
    int32_t *dst;           // this lane's column of the channel's [frame_length][32] tile
    __device__ __forceinline__ void operator()(uint32_t j, int32_t v) { dst[(size_t)j * 32u] = v; }
};

template <int DEPTH>
__global__ void __launch_bounds__(kRingStride) dec_lane_kernel(DecArgs A)
{
    __shared__ uint32_t s_ring[kRingSlots][kRingStride];
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= A.num_packets) return;
    const uint32_t pkt = A.perm[A.pkt_base + tid];
    const uint32_t nch = A.num_channels;
    const uint32_t size = A.pkt_size[pkt];
    const uint32_t cap_bits = size * 8u;
    const uint32_t slot_samples = A.pkt_samples[pkt];
    const uint32_t F = A.frame_length;
    int32_t *tile0 = A.chan_scratch + ((size_t)(tid >> 5) * nch * F) * 32u + (tid & 31u);
    DecChanMeta *meta = A.chan_meta + (size_t)pkt * nch;
    for (uint32_t c = 0; c < nch; c++) {            // channels that never arrive are zero-filled (:972-998)
        DecChanMeta z;
        z.n = slot_samples; z.shift_pos = 0; z.kind = CH_ZERO; z.shift = 0; z.mix_res = 0; z.mix_bits = 0;
        meta[c] = z;
    }
    BitReader br;
    br.start(A.packets + A.pkt_off[pkt], size, &s_ring[0][threadIdx.x]);

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
            br.pos += 4;                                                        // element instance tag
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
            ScratchOut ou, ov;
            ou.dst = tile0 + (size_t)channel_index * F * 32u;
            ov.dst = ou.dst + (size_t)F * 32u;
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
                status = decode_channel(br, cap_bits, A, n, chan_bits, hu, ou);
                if (status == 0 && pair) status = decode_channel(br, cap_bits, A, n, chan_bits, hv, ov);
            } else {
                // uncompressed element (:697-727, :856-896): raw samples, pairs interleaved
                const uint32_t sh = 32u - DEPTH;
                for (uint32_t j = 0; j < n; j++) {
                    ou(j, (int32_t)(br.get(DEPTH) << sh) >> sh);
                    if (pair) ov(j, (int32_t)(br.get(DEPTH) << sh) >> sh);
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
    A.pkt_status[pkt] = status;
}

// grid: x = 32-sample tiles of a frame, y = groups of 32 packets (permuted order), z = channel
template <int DEPTH>
__global__ void __launch_bounds__(256) dec_output_kernel(DecArgs A)
{
    __shared__ int32_t su[32][33], sv[32][33];
    constexpr uint32_t bps = DepthTraits<DEPTH>::kBytes;
    const uint32_t nch = A.num_channels, F = A.frame_length;
    const uint32_t group = blockIdx.y, c = blockIdx.z, j0 = blockIdx.x * 32u;
    const uint32_t tx = threadIdx.x, ty = threadIdx.y;
    const uint32_t stride = nch * bps;

    // does any packet of the group need this channel slot written from here? (CH_PAIR_V is written by its U)
    const uint32_t my_slot = group * 32u + tx;
    uint32_t kind_x = CH_PAIR_V;
    if (my_slot < A.num_packets && ty == 0) {
        const DecChanMeta &m = A.chan_meta[(size_t)A.perm[A.pkt_base + my_slot] * nch + c];
        kind_x = (j0 < m.n) ? m.kind : (uint32_t)CH_PAIR_V;
    }
    const int any = __syncthreads_or(kind_x != CH_PAIR_V);
    if (!any) return;

    const int32_t *tile_u = A.chan_scratch + ((size_t)(group * nch + c) * F) * 32u;
    const int32_t *tile_v = tile_u + (size_t)F * 32u;
    const bool have_v = (c + 1 < nch);
    for (uint32_t r = ty; r < 32; r += 8) {
        const uint32_t j = j0 + r;
        su[r][tx] = j < F ? tile_u[(size_t)j * 32u + tx] : 0;
        sv[r][tx] = (have_v && j < F) ? tile_v[(size_t)j * 32u + tx] : 0;
    }
    __syncthreads();

    const uint32_t j = j0 + tx;
    for (uint32_t pr = ty; pr < 32; pr += 8) {
        const uint32_t slot = group * 32u + pr;
        if (slot >= A.num_packets) break;
        const uint32_t pkt = A.perm[A.pkt_base + slot];
        const DecChanMeta m = A.chan_meta[(size_t)pkt * nch + c];
        if (j >= m.n || m.kind == CH_PAIR_V) continue;
        uint8_t *out = A.pcm_out + (A.out_frame[pkt] + j) * stride + (size_t)c * bps;
        int32_t l = su[tx][pr];
        if (m.kind == CH_ZERO) {
            store_sample<DEPTH>(out, 0);
            continue;
        }
        BitPeek bp;
        if (m.shift) bp.start(A.packets + A.pkt_off[pkt], A.pkt_size[pkt]);
        if (m.kind == CH_MONO) {
            if (m.shift) l = (int32_t)(((uint32_t)l << m.shift) | bp.bits_at(m.shift_pos + j * m.shift, m.shift));   // :436-495
            store_sample<DEPTH>(out, l);
        } else {
            const int32_t v = sv[tx][pr];
            int32_t r;
            if (m.mix_res != 0) {                       // :193-223
                l = l + v - (((int32_t)m.mix_res * v) >> m.mix_bits);
                r = l - v;
            } else {
                r = v;
            }
            if (m.shift) {                              // :282-383
                const uint32_t both = bp.bits_at(m.shift_pos + j * 2u * m.shift, 2u * m.shift);
                l = (int32_t)(((uint32_t)l << m.shift) | (both >> m.shift));
                r = (int32_t)(((uint32_t)r << m.shift) | (both & ((1u << m.shift) - 1u)));
            }
            store_sample<DEPTH>(out, l);
            store_sample<DEPTH>(out + bps, r);
        }
    }
}


// ---- CAF packet table on the device -------------------------------------------------------
// The 'pakt' chunk stores each packet size as a BER integer (7 bits per byte, high bit = "more",
// convert-utility/CAFFileALAC.cpp:189-258).  ber_flag_kernel marks the last byte of every entry; an exclusive
// scan of the marks numbers the entries; ber_value_kernel assembles each entry from the <= 5 bytes that end at
// its mark; ber_count_kernel finds where the reference's decode loop would stop (a zero size, or a packet that
// no longer fits the data chunk, convert-utility/main.cu:717).
__global__ void ber_flag_kernel(const uint8_t *table, uint64_t nbytes, uint32_t *flags)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nbytes) flags[i] = (table[i] & 0x80u) ? 0u : 1u;
}

__global__ void ber_value_kernel(const uint8_t *table, uint64_t nbytes, const uint64_t *entry_of_byte, uint32_t *sizes, uint64_t cap)
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

__global__ void ber_count_kernel(const uint32_t *sizes, const uint64_t *offsets, uint64_t entries, uint64_t data_bytes,
                                 unsigned long long *first_bad)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= entries) return;
    if (sizes[k] == 0 || offsets[k] + sizes[k] > data_bytes) atomicMin(first_bad, (unsigned long long)k);
}

}  // namespace alacb
