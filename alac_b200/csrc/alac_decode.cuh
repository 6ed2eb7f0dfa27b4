// alac_decode.cuh -- decode kernels.
//
//   dec_header_kernel   one lane per packet: reads the first audio element's header to learn the
//                       packet's sample count (partial-frame field), so output offsets can be scanned.
//   dec_packet_kernel   one lane per packet: ALACDecoder::Decode (codec/ALACDecoder.cu:571-1002)
//                       with dyn_decomp, unpc_block, the shift-byte merge and unmixNN /
//                       copyPredictorToNN (codec/ALACDecoder.cu:193-495) fused per sample.
//                       A pair's U samples are parked in the pair's own output slots (>= 4 bytes
//                       per sample-frame) until V arrives, so no side buffer exists.
#pragma once
#include "alac_device.cuh"

namespace alacb {

struct DecArgs {
    const uint8_t *packets;
    const uint64_t *pkt_off;      // byte offset of each packet (exclusive scan of sizes)
    const uint32_t *pkt_size;
    uint32_t num_packets;
    uint32_t frame_length, pb, mb, kb, num_channels;
    uint8_t *pcm_out;
    const uint64_t *out_frame;    // first output sample-frame of each packet (exclusive scan)
    uint32_t *pkt_samples;
    int32_t *pkt_status;
};

// walk element tags until the first SCE/LFE/CPE and return its sample count
__global__ void dec_header_kernel(DecArgs A)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= A.num_packets) return;
    const uint32_t size = A.pkt_size[p];
    BitReader br;
    br.start(A.packets + A.pkt_off[p], size);
    uint32_t n = A.frame_length;
    for (int guard = 0; guard < 64; guard++) {
        if (!((br.pos >> 3) < size)) break;
        const uint32_t tag = br.get(3);
        if (tag == ID_SCE || tag == ID_LFE || tag == ID_CPE) {
            br.pos += 4 + 12;
            const uint32_t hb = br.get(4);
            if (hb >> 3) { n = br.get(16) << 16; n |= br.get(16); }
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
}

// U samples parked in the pair's output slot (2 * bytes-per-sample >= 4 bytes, see file header)
template <int DEPTH>
__device__ __forceinline__ void park_store(uint8_t *p, int32_t v)
{
    if (DEPTH == 16) {
        reinterpret_cast<uint16_t *>(p)[0] = (uint16_t)v;
        reinterpret_cast<uint16_t *>(p)[1] = (uint16_t)((uint32_t)v >> 16);
    } else if (DEPTH == 32) {
        *reinterpret_cast<int32_t *>(p) = v;
    } else {
        p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16);
    }
}
template <int DEPTH>
__device__ __forceinline__ int32_t park_load(const uint8_t *p)
{
    if (DEPTH == 16) {
        return (int32_t)((uint32_t)reinterpret_cast<const uint16_t *>(p)[0] | ((uint32_t)reinterpret_cast<const uint16_t *>(p)[1] << 16));
    } else if (DEPTH == 32) {
        return *reinterpret_cast<const int32_t *>(p);
    } else {
        const uint32_t w = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
        return (int32_t)(w << 8) >> 8;
    }
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
template <int TAPS, class Out>
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
        out(j, predict_dec_step<TAPS, true>(r, hist, a, chanshift));
    }
}

// any numactive 0..31, any denShift, mode != 0 (codec/dp_dec.c:67-95, :335-380; codec/ALACDecoder.cu:686-694)
template <class Out>
__device__ __noinline__ void decode_channel_general(BitReader &br, uint32_t cap_bits, AgDec &ag, uint32_t n,
                                                    ChanHeader &h, uint32_t chanshift, Out &out)
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
}

template <class Out>
__device__ __forceinline__ int32_t decode_channel(BitReader &br, uint32_t cap_bits, const DecArgs &A, uint32_t n,
                                                  uint32_t chan_bits, ChanHeader &h, Out &out)
{
    AgDec ag;
    ag.start(br, n, A.mb, (A.pb * h.pb_factor) / 4, A.kb, chan_bits);       // codec/ALACDecoder.cu:682
    const uint32_t chanshift = 32u - chan_bits;
    if (h.mode == 0 && h.den_shift == kDenShift && h.num == 4) decode_channel_fast<4>(br, cap_bits, ag, n, h, chanshift, out);
    else if (h.mode == 0 && h.den_shift == kDenShift && h.num == 8) decode_channel_fast<8>(br, cap_bits, ag, n, h, chanshift, out);
    else decode_channel_general(br, cap_bits, ag, n, h, chanshift, out);
    // dyn_decomp's exit check "cur <= end" (codec/ag_dec.c:359)
    if (!ag.status && (br.pos >> 3) > (cap_bits >> 3)) ag.status = -50;
    return ag.status;
}

template <int DEPTH>
struct MonoOut {
    uint8_t *base; uint32_t stride; uint32_t shift; BitReader sr;
    __device__ __forceinline__ void operator()(uint32_t j, int32_t v)
    {
        if (shift) v = (int32_t)(((uint32_t)v << shift) | sr.get(shift));     // codec/ALACDecoder.cu:436-495
        store_sample<DEPTH>(base + (size_t)j * stride, v);
    }
};
template <int DEPTH>
struct ParkOut {
    uint8_t *base; uint32_t stride;
    __device__ __forceinline__ void operator()(uint32_t j, int32_t v) { park_store<DEPTH>(base + (size_t)j * stride, v); }
};
template <int DEPTH>
struct PairOut {
    uint8_t *base; uint32_t stride; uint32_t shift; int32_t mix_res; uint32_t mix_bits; BitReader sr;
    __device__ __forceinline__ void operator()(uint32_t j, int32_t v)
    {
        uint8_t *p = base + (size_t)j * stride;
        const int32_t u = park_load<DEPTH>(p);
        int32_t l, r;
        if (mix_res != 0) {                         // codec/ALACDecoder.cu:193-223
            l = u + v - ((mix_res * v) >> mix_bits);
            r = l - v;
        } else {
            l = u;
            r = v;
        }
        if (shift) {                                // :282-383
            const uint32_t both = sr.get(2 * shift);
            l = (int32_t)(((uint32_t)l << shift) | (both >> shift));
            r = (int32_t)(((uint32_t)r << shift) | (both & ((1u << shift) - 1u)));
        }
        store_sample<DEPTH>(p, l);
        store_sample<DEPTH>(p + DepthTraits<DEPTH>::kBytes, r);
    }
};

template <int DEPTH>
__global__ void __launch_bounds__(128) dec_packet_kernel(DecArgs A)
{
    const uint32_t pkt = blockIdx.x * blockDim.x + threadIdx.x;
    if (pkt >= A.num_packets) return;
    constexpr uint32_t bps = DepthTraits<DEPTH>::kBytes;
    const uint32_t nch = A.num_channels;
    const uint32_t stride = nch * bps;
    const uint32_t size = A.pkt_size[pkt];
    const uint32_t cap_bits = size * 8u;
    const uint32_t slot_samples = A.pkt_samples[pkt];
    uint8_t *out_base = A.pcm_out + A.out_frame[pkt] * stride;
    BitReader br;
    br.start(A.packets + A.pkt_off[pkt], size);

    uint32_t n = A.frame_length;
    uint32_t channel_index = 0;
    int32_t status = 0;
    ChanHeader hu, hv;

    while (status == 0) {
        if (!((br.pos >> 3) < size)) { status = -50; break; }                   // :615
        const uint32_t tag = br.get(3);
        if (tag == ID_SCE || tag == ID_LFE) {
            br.pos += 4;                                                        // element instance tag
            if (br.get(12) != 0) { status = -50; break; }                       // :633
            const uint32_t hb = br.get(4);
            const uint32_t partial = hb >> 3;
            uint32_t bytes_shifted = (hb >> 1) & 3u;
            if (bytes_shifted == 3) { status = -50; break; }                    // :641
            const uint32_t escape = hb & 1u;
            const uint32_t chan_bits = DEPTH - bytes_shifted * 8;
            if (partial) { n = br.get(16) << 16; n |= br.get(16); }             // :650-654
            if (n > slot_samples) { status = -50; break; }
            const bool in_range = channel_index < nch;
            MonoOut<DEPTH> out;
            out.base = out_base + (size_t)(in_range ? channel_index : 0) * bps;
            out.stride = stride;
            out.shift = 0;
            if (!in_range) { status = -50; break; }
            if (!escape) {
                br.pos += 16;                                                   // mixBits, mixRes
                read_chan_header(br, hu);
                if (bytes_shifted) {                                            // :675-679
                    out.shift = bytes_shifted * 8;
                    out.sr = br;
                    br.pos += out.shift * n;
                }
                status = decode_channel(br, cap_bits, A, n, chan_bits, hu, out);
            } else {
                const uint32_t sh = 32u - chan_bits;                            // :697-727
                for (uint32_t j = 0; j < n; j++) out(j, (int32_t)(br.get(chan_bits) << sh) >> sh);
            }
            channel_index += 1;
        } else if (tag == ID_CPE) {
            if (channel_index + 2 > nch) break;                                 // :759-760
            br.pos += 4;
            if (br.get(12) != 0) { status = -50; break; }
            const uint32_t hb = br.get(4);
            const uint32_t partial = hb >> 3;
            const uint32_t bytes_shifted = (hb >> 1) & 3u;
            if (bytes_shifted == 3) { status = -50; break; }
            const uint32_t escape = hb & 1u;
            const uint32_t chan_bits = DEPTH - bytes_shifted * 8 + 1;
            if (partial) { n = br.get(16) << 16; n |= br.get(16); }
            if (n > slot_samples) { status = -50; break; }
            uint8_t *eb = out_base + (size_t)channel_index * bps;
            PairOut<DEPTH> pout;
            pout.base = eb; pout.stride = stride; pout.shift = 0; pout.mix_res = 0; pout.mix_bits = 0;
            if (!escape) {
                pout.mix_bits = br.get(8);
                pout.mix_res = (int32_t)(int8_t)br.get(8);
                read_chan_header(br, hu);
                read_chan_header(br, hv);
                if (bytes_shifted) {                                            // :818-822
                    pout.shift = bytes_shifted * 8;
                    pout.sr = br;
                    br.pos += pout.shift * 2 * n;
                }
                ParkOut<DEPTH> park;
                park.base = eb; park.stride = stride;
                status = decode_channel(br, cap_bits, A, n, chan_bits, hu, park);
                if (status) break;
                status = decode_channel(br, cap_bits, A, n, chan_bits, hv, pout);
            } else {
                const uint32_t sh = 32u - DEPTH;                                // :856-896
                for (uint32_t j = 0; j < n; j++) {
                    const int32_t l = (int32_t)(br.get(DEPTH) << sh) >> sh;
                    const int32_t r = (int32_t)(br.get(DEPTH) << sh) >> sh;
                    uint8_t *p = eb + (size_t)j * stride;
                    store_sample<DEPTH>(p, l);
                    store_sample<DEPTH>(p + bps, r);
                }
            }
            channel_index += 2;
        } else if (tag == ID_CCE || tag == ID_PCE) {
            status = -50;                                                       // :932-939
        } else if (tag == ID_DSE) {
            br.pos += 4;
            const uint32_t align = br.get(1);
            uint32_t count = br.get(8);
            if (count == 255) count += br.get(8);
            if (align && (br.pos & 7u)) br.pos += 8u - (br.pos & 7u);
            br.pos += count * 8;
            if ((br.pos >> 3) > size) status = -50;
        } else if (tag == ID_FIL) {
            int32_t count = (int32_t)br.get(4);
            if (count == 15) count += (int32_t)br.get(8) - 1;
            br.pos += (uint32_t)count * 8;
            if ((br.pos >> 3) > size) status = -50;
        } else {
            break;                                                              // ID_END :955-961
        }
        if (channel_index >= nch) break;                                        // :966-967
    }
    // channels that never arrived are zero-filled (:972-998)
    if (status == 0) {
        for (; channel_index < nch; channel_index++)
            for (uint32_t j = 0; j < slot_samples; j++)
                store_sample<DEPTH>(out_base + ((size_t)j * nch + channel_index) * bps, 0);
    }
    A.pkt_status[pkt] = status;
}

}  // namespace alacb
