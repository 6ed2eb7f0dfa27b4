// alac_encode.cuh -- encode kernels.
//
//   enc_search_split_kernel + enc_final2_kernel / enc_final_kernel   (frames_per_segment = 1: every frame is its own chain)
//                       one lane per (frame, channel), one-warp CTAs.  The search kernel runs stages A and B of
//                       EncodeStereo / EncodeMono (codec/ALACEncoder.cu:290-558, :812-963) -- mixRes search,
//                       numU/numV search, escape estimate; U and V of a pair sit on adjacent lanes and exchange
//                       bit counts by shuffle -- and files each channel's final-pass job under its tap count; the
//                       final kernel runs stage C (final predictor + Golomb emission) with full, uniform warps.
//                       Dense streams (mono / stereo, packets on 16-byte boundaries) stream their PCM through the
//                       block ring (QuadRing: 16-byte cp.async, three blocks ahead): enc_search_split_kernel<.., DENSE>
//                       and enc_final2_kernel -- the latter as a one-warp form, or as a two-warp form (predictor warp ->
//                       residual tiles in shared memory -> Golomb warp) for launches that do not fill the GPU.
//                       Everything else goes through enc_final_kernel (per-sample loads).
//   enc_search_kernel   chained frames (frames_per_segment != 1 or a coefficient-state hand-off): one lane per
//                       (segment, channel) walks the frames of its segment; stages A, B, C in one kernel, the
//                       final-pass jobs regrouped by tap count inside the CTA.
//                       All forms emit each channel's Golomb stream into a private scratch slab and one
//                       ElemRec per element.
//   enc_size_kernel     packet byte sizes from the element records.
//   scan_*_kernel       exclusive scan of the sizes (two launches, chained across chunks).
//   xchg_*_kernel       the cross-GPU packet-offset exchange and the gap compaction of staged placement.
//   enc_assemble_kernel one warp per packet: gathers header / shift bytes / Golomb streams /
//                       escape samples into the packet at its scanned offset (32-bit stores); words inside one long
//                       region take a short path (funnel shift of two slab words, PCM span staged in shared memory).
#pragma once
#include "alac_device.cuh"
#include <type_traits>

namespace alacb {

// per (packet, element) result of the search kernel
struct ElemRec {
    uint32_t bits_u, bits_v;      // final-pass Golomb bits per channel
    uint32_t elem_bits;           // total element size in bits incl. tag + instance
    uint8_t escape;               // 0 compressed, 1 escape by estimate, 2 escape by post-check
    uint8_t mix_res, num_u, num_v;
    int16_t coef_u[8], coef_v[8]; // coefficients as written in the header (before the final pass)
};

struct EncLayout {
    uint32_t channels;
    uint32_t frame_size;
    uint32_t fast_mode;
    uint32_t elems_per_packet;
    uint32_t chains_per_packet;
    // per element (packet order): tag, first channel, per-type instance tag, first chain index
    uint8_t elem_tag[8], elem_chan[8], elem_inst[8], elem_chain[8];
};

struct EncArgs {
    const uint8_t *pcm;           // device, interleaved
    const uint64_t *pkt_frame;    // per packet: first sample-frame (index into pcm)
    const uint32_t *pkt_samples;  // per packet: sample-frames in it
    const uint32_t *seg_first;    // per segment: first packet
    const uint32_t *seg_count;    // per segment: packets in it
    const uint32_t *seg_stream;   // per segment: stream index | 0x80000000 if first | 0x40000000 if last of stream
    uint32_t seg_base;            // chunk: first segment handled by this launch
    uint32_t num_segments;        // chunk: segments handled by this launch
    uint32_t pkt_base;            // chunk: first packet (recs / scratch are chunk-relative)
    EncLayout lay;
    ElemRec *recs;                // [packet][elem]
    uint32_t *scratch;            // [packet][chain][cap_words]
    uint32_t cap_words;
    int16_t *state;               // optional [stream][8][2][2][8]
};

// ---- sample sources -------------------------------------------------------------------------------
// One lane's view of its channel after the reference's mix/copy stage (codec/matrix_enc.cu,
// codec/ALACEncoder.cu:1144-1382): get(j) is u[j] (U lane) or v[j] (V lane); samples at or
// beyond `valid` read as zero (deterministic-padding rule).
// Loading is split from mixing so the predictor loop can issue the load of sample j + kAhead
// while it works on sample j: fetch() only loads (no arithmetic on the result), mix() unpacks.
// PACKED: a pure stereo stream (L, R adjacent, nothing in between) at natural alignment, so one
// sample-frame is a single 32/64-bit load (16/32-bit) or three 16-bit loads (20/24-bit).
// PCM ring of the wide path: per lane kPcmSlots slots of 16 bytes (four 16-bit stereo sample-frames), slot s of lane l
// at [s][l], so a warp's 128-bit accesses are conflict-free.  cp.async fills it kPcmSlots - 1 blocks ahead of the
// arithmetic; unlike a register queue (whose rotating moves wait for the newest load once per block) nothing waits
// on a load before its data is due.
constexpr uint32_t kPcmSlots = 4;
template <uint32_t THREADS> struct PcmRing { uint4 slot[kPcmSlots][THREADS]; };
template <uint32_t THREADS> __device__ __forceinline__ uint32_t pcm_ring_addr(PcmRing<THREADS> &r)
{
    return (uint32_t)__cvta_generic_to_shared(&r.slot[0][threadIdx.x]);
}
__device__ __forceinline__ uint4 lds_u128(uint32_t smem_addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_addr) : "memory");
    return v;
}
__device__ __forceinline__ void cp_async_u128(uint32_t smem_dst, const void *gsrc, uint32_t src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}

template <int DEPTH, bool STEREO, bool PACKED>
struct MixSrc {
    static constexpr int kWords = !STEREO ? ((DEPTH == 16 || DEPTH == 32) ? 1 : 3)
                                  : PACKED ? (DEPTH == 16 ? 1 : DEPTH == 32 ? 2 : 3)
                                           : ((DEPTH == 16 || DEPTH == 32) ? 2 : 6);
    static constexpr int kAhead = kWords <= 2 ? 4 : 2;      // prefetch distance in samples
    // 16-bit packed stereo: the main loop of predict_pass reads four sample-frames per 16-byte load
    static constexpr bool kWide = STEREO && PACKED && DEPTH == 16;
    static constexpr bool kDense = false;
    struct Raw { uint32_t w[kWords]; };

    const uint8_t *base;    // sample-frame 0 of the packet, first channel of the element
    uint32_t stride;        // bytes per sample-frame
    uint32_t ring;          // kWide: shared-memory address of this lane's slot 0 of the PCM ring
    uint32_t ring_stride;   // kWide: bytes from one slot to the next (16 x threads of the CTA)
    uint32_t valid;
    int32_t mix_res;
    bool is_v;

    __device__ __forceinline__ Raw fetch(uint32_t j) const
    {
        Raw r;
        const uint8_t *p = base + (size_t)j * stride;
        if (!STEREO) {
            if (DEPTH == 16) r.w[0] = __ldg(reinterpret_cast<const uint16_t *>(p));
            else if (DEPTH == 32) r.w[0] = __ldg(reinterpret_cast<const uint32_t *>(p));
            else { r.w[0] = __ldg(p); r.w[1] = __ldg(p + 1); r.w[2] = __ldg(p + 2); }
        } else if (PACKED) {
            if (DEPTH == 16) r.w[0] = __ldg(reinterpret_cast<const uint32_t *>(p));
            else if (DEPTH == 32) { const uint2 t = __ldg(reinterpret_cast<const uint2 *>(p)); r.w[0] = t.x; r.w[1] = t.y; }
            else { const uint16_t *q = reinterpret_cast<const uint16_t *>(p); r.w[0] = __ldg(q); r.w[1] = __ldg(q + 1); r.w[2] = __ldg(q + 2); }
        } else {
            if (DEPTH == 16) { r.w[0] = __ldg(reinterpret_cast<const uint16_t *>(p)); r.w[1] = __ldg(reinterpret_cast<const uint16_t *>(p + 2)); }
            else if (DEPTH == 32) { r.w[0] = __ldg(reinterpret_cast<const uint32_t *>(p)); r.w[1] = __ldg(reinterpret_cast<const uint32_t *>(p + 4)); }
            else { for (int i = 0; i < 6; i++) r.w[i] = __ldg(p + i); }
        }
        return r;
    }
    static __device__ __forceinline__ int32_t widen(uint32_t w24)   // 20/24-bit container -> sample >> shift
    {
        return DEPTH == 20 ? ((int32_t)(w24 << 8) >> 12) : ((int32_t)(w24 << 8) >> (8 + DepthTraits<DEPTH>::kShift));
    }
    __device__ __forceinline__ int32_t mix(const Raw &r) const
    {
        constexpr uint32_t sh = DepthTraits<DEPTH>::kShift;
        int32_t l, rr;
        if (!STEREO) {
            if (DEPTH == 16) return (int32_t)(int16_t)r.w[0];
            if (DEPTH == 32) return (int32_t)r.w[0] >> sh;
            return widen(r.w[0] | (r.w[1] << 8) | (r.w[2] << 16));
        } else if (PACKED) {
            if (DEPTH == 16) { l = (int32_t)(int16_t)(r.w[0] & 0xffffu); rr = (int32_t)r.w[0] >> 16; }
            else if (DEPTH == 32) { l = (int32_t)r.w[0] >> sh; rr = (int32_t)r.w[1] >> sh; }
            else { l = widen(r.w[0] | ((r.w[1] & 0xffu) << 16)); rr = widen((r.w[1] >> 8) | (r.w[2] << 8)); }
        } else {
            if (DEPTH == 16) { l = (int32_t)(int16_t)r.w[0]; rr = (int32_t)(int16_t)r.w[1]; }
            else if (DEPTH == 32) { l = (int32_t)r.w[0] >> sh; rr = (int32_t)r.w[1] >> sh; }
            else { l = widen(r.w[0] | (r.w[1] << 8) | (r.w[2] << 16)); rr = widen(r.w[3] | (r.w[4] << 8) | (r.w[5] << 16)); }
        }
        // u = (mixRes*l + (4-mixRes)*r) >> 2, v = l - r; mixRes 0: u = l, v = r (codec/matrix_enc.cu:72-99).
        // All four cases are (cl*l + cr*r) >> sh with per-lane constants, so the hot loop has no selects.
        return (cl * l + cr * rr) >> sh_mix;
    }
    int32_t cl, cr;
    uint32_t sh_mix;
    __device__ __forceinline__ void set_mix(int32_t res, bool v)
    {
        mix_res = res;
        is_v = v;
        if (res != 0) { cl = v ? 1 : res; cr = v ? -1 : (1 << kMixBits) - res; sh_mix = v ? 0u : (uint32_t)kMixBits; }
        else { cl = v ? 0 : 1; cr = v ? 1 : 0; sh_mix = 0u; }
    }
    __device__ __forceinline__ int32_t get(uint32_t j) const
    {
        if (j >= valid) return 0;
        return mix(fetch(j));
    }
};

// ---- dense elements: the word ring (used by the dense search passes and by enc_final2_kernel, see there) ----------
template <int DEPTH, bool STEREO> struct DenseElem {
    static constexpr uint32_t kFrameBytes = DepthTraits<DEPTH>::kBytes * (STEREO ? 2u : 1u);
    static constexpr uint32_t kQuadWords = kFrameBytes;          // 4 sample-frames = kFrameBytes 32-bit words
};
#ifndef ALAC_QUAD_AHEAD
#define ALAC_QUAD_AHEAD 3
#endif
constexpr uint32_t kQuadAhead = ALAC_QUAD_AHEAD;
constexpr uint32_t kEncTileRows = 32;

// sign-extended 16-bit value at byte offset o (0..3) + 1 of the word pair (lo, hi): bytes o+1, o+2 -> (b2 << 8 | b1)
template <uint32_t O> __device__ __forceinline__ int32_t prmt_s16_at(uint32_t lo, uint32_t hi)
{
    constexpr uint32_t b1 = O + 1u, b2 = O + 2u;                 // byte indices into {lo: 0..3, hi: 4..7}
    constexpr uint32_t sel = b1 | (b2 << 4) | ((b2 | 8u) << 8) | ((b2 | 8u) << 12);     // nibble msb = replicate the byte's sign
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(hi), "n"(sel));       // (__byte_perm ignores the msb of a nibble)
    return (int32_t)d;
}
// 24-bit container at byte offset O of the word pair, as a left-justified 32-bit word (sample << 8)
template <uint32_t O> __device__ __forceinline__ uint32_t prmt_u24hi_at(uint32_t lo, uint32_t hi)
{
    constexpr uint32_t sel = (O << 4) | ((O + 1u) << 8) | ((O + 2u) << 12);   // byte 0 <- lo.b0 (garbage, shifted out below)
    return __byte_perm(lo, hi, sel) & 0xffffff00u;
}

// the four (left, right) -- or four mono -- samples of a quad, after the depth's shift (what the predictor sees)
template <int DEPTH, bool STEREO>
__device__ __forceinline__ void unpack_quad(const uint32_t (&w)[DenseElem<DEPTH, STEREO>::kQuadWords], int32_t (&l)[4], int32_t (&r)[4])
{
    constexpr uint32_t sh = DepthTraits<DEPTH>::kShift;
    if (DEPTH == 16) {
        if (STEREO) {
#pragma unroll
            for (int i = 0; i < 4; i++) { l[i] = (int32_t)(int16_t)(w[i] & 0xffffu); r[i] = (int32_t)w[i] >> 16; }
        } else {
            l[0] = (int32_t)(int16_t)(w[0] & 0xffffu); l[1] = (int32_t)w[0] >> 16;
            l[2] = (int32_t)(int16_t)(w[1] & 0xffffu); l[3] = (int32_t)w[1] >> 16;
        }
    } else if (DEPTH == 32) {
        if (STEREO) {
#pragma unroll
            for (int i = 0; i < 4; i++) { l[i] = (int32_t)w[2 * i] >> sh; r[i] = (int32_t)w[2 * i + 1] >> sh; }
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) l[i] = (int32_t)w[i] >> sh;
        }
    } else if (DEPTH == 24) {
        // sample >> 8 = the sign-extended 16-bit value at byte offset + 1
        if (STEREO) {           // L at bytes 0, 6, 12, 18; R at 3, 9, 15, 21 of the quad's 6 words
            l[0] = prmt_s16_at<0>(w[0], w[1]); r[0] = prmt_s16_at<3>(w[0], w[1]);
            l[1] = prmt_s16_at<2>(w[1], w[2]); r[1] = prmt_s16_at<1>(w[2], w[3]);
            l[2] = prmt_s16_at<0>(w[3], w[4]); r[2] = prmt_s16_at<3>(w[3], w[4]);
            l[3] = prmt_s16_at<2>(w[4], w[5]); r[3] = prmt_s16_at<1>(w[5], w[5]);
        } else {                // samples at bytes 0, 3, 6, 9 of the quad's 3 words
            l[0] = prmt_s16_at<0>(w[0], w[1]); l[1] = prmt_s16_at<3>(w[0], w[1]);
            l[2] = prmt_s16_at<2>(w[1], w[2]); l[3] = prmt_s16_at<1>(w[2], w[2]);
        }
    } else {                    // 20-bit, left-justified in 3 bytes: (w24 << 8) >> 12
        if (STEREO) {
            l[0] = (int32_t)prmt_u24hi_at<0>(w[0], w[1]) >> 12; r[0] = (int32_t)prmt_u24hi_at<3>(w[0], w[1]) >> 12;
            l[1] = (int32_t)prmt_u24hi_at<2>(w[1], w[2]) >> 12; r[1] = (int32_t)prmt_u24hi_at<1>(w[2], w[3]) >> 12;
            l[2] = (int32_t)prmt_u24hi_at<0>(w[3], w[4]) >> 12; r[2] = (int32_t)prmt_u24hi_at<3>(w[3], w[4]) >> 12;
            l[3] = (int32_t)prmt_u24hi_at<2>(w[4], w[5]) >> 12; r[3] = (int32_t)prmt_u24hi_at<1>(w[5], w[5]) >> 12;
        } else {
            l[0] = (int32_t)prmt_u24hi_at<0>(w[0], w[1]) >> 12; l[1] = (int32_t)prmt_u24hi_at<3>(w[0], w[1]) >> 12;
            l[2] = (int32_t)prmt_u24hi_at<2>(w[1], w[2]) >> 12; l[3] = (int32_t)prmt_u24hi_at<1>(w[2], w[2]) >> 12;
        }
    }
}

// One lane's ring over its packet's PCM.  The unit of transfer is a BLOCK: the smallest run of whole quads whose size
// is a multiple of 16 bytes (16-bit stereo: 1 quad = 16 B; 24-bit stereo: 2 quads = 48 B; 32-bit stereo: 1 quad = 32 B;
// mono: 2, 4, 1 quads), copied by 16-byte cp.async -- a warp's lanes sit in 32 different packets, so every request
// costs 32 L1 wavefronts whatever its size, and the encode kernels of the wider depths were bound by exactly that
// (l1tex 73-87 % busy with 4-byte requests).  kQuadAhead + 1 slots of one block each; granule g of a slot at
// [slot * G + g][lane] as uint4, so the 128-bit accesses of a warp are conflict-free.  Only whole blocks of the packet
// are ever read (the odd frames at either end go through the scalar path); a request past them copies nothing.
// Needs the packet's first byte on a 16-byte boundary.
template <int DEPTH, bool STEREO>
struct QuadRing {
    static constexpr uint32_t WQ = DenseElem<DEPTH, STEREO>::kQuadWords;
    static constexpr uint32_t QB = (WQ % 4u == 0) ? 1u : (WQ % 2u == 0) ? 2u : 4u;     // quads per block
    static constexpr uint32_t G = WQ * QB / 4u;                                       // 16-byte granules per block
    static constexpr uint32_t kBlockFrames = 4u * QB;
    static constexpr uint32_t kSlots = kQuadAhead + 1u;
    static constexpr uint32_t kRows = G * kSlots;       // the CTA's ring array is uint4[kRows][32]
    static_assert((kSlots & (kSlots - 1u)) == 0, "slot index is a mask");
    const uint4 *base;          // granule 0 = sample-frame 0 of the packet (16-byte aligned)
    uint32_t ring;              // shared-memory address of this lane's column
    uint32_t blocks;            // whole blocks in the packet
    __device__ __forceinline__ void start(const uint8_t *packet, uint32_t n, uint4 *ring_column)
    {
        base = reinterpret_cast<const uint4 *>(packet);
        ring = (uint32_t)__cvta_generic_to_shared(ring_column);
        blocks = n / kBlockFrames;
    }
    __device__ __forceinline__ uint32_t slot(uint32_t b) const { return ring + (b & (kSlots - 1u)) * (G * 512u); }
    __device__ __forceinline__ void request(uint32_t b) const
    {
        if (b < blocks) {
            const uint32_t dst = slot(b);
            const uint4 *g = base + (size_t)b * G;
#pragma unroll
            for (uint32_t i = 0; i < G; i++)
#ifndef ALAC_RING_CG
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + i * 512u), "l"(g + i) : "memory");
#else
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + i * 512u), "l"(g + i) : "memory");
#endif
        }
        cp_async_commit();
    }
    __device__ __forceinline__ void read(uint32_t b, uint32_t (&w)[WQ * QB]) const
    {
        const uint32_t src = slot(b);
#pragma unroll
        for (uint32_t i = 0; i < G; i++) {
            const uint4 v = lds_u128(src + i * 512u);
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
        }
    }
};

#ifdef ALAC_RING_BULK
// EXPERIMENT: the same ring filled by per-lane bulk copies (cp.async.bulk, SASS UBLKCP) that complete on an
// mbarrier per slot instead of 16-byte cp.async groups.  A lane's slots are contiguous (a bulk copy writes one run of
// shared memory): row = kSlots blocks + 16 bytes of padding, which keeps the 128-bit reads of a quarter-warp on
// different banks.  Every lane arrives once per block index on the slot's barrier (count 32), with its byte count when
// it has a block to fetch and without when it has not, so all lanes of the warp must walk the same block indices.
template <int DEPTH, bool STEREO>
struct BulkRing {
    static constexpr uint32_t WQ = DenseElem<DEPTH, STEREO>::kQuadWords;
    static constexpr uint32_t QB = (WQ % 4u == 0) ? 1u : (WQ % 2u == 0) ? 2u : 4u;
    static constexpr uint32_t G = WQ * QB / 4u;
    static constexpr uint32_t kBlockFrames = 4u * QB;
    static constexpr uint32_t kSlots = kQuadAhead + 1u;
    static constexpr uint32_t kRowBytes = kSlots * G * 16u + 16u;
    const uint8_t *base;
    uint32_t row, bars, blocks, first;
    __device__ __forceinline__ void start(const uint8_t *packet, uint32_t n, uint8_t *row_ptr, uint64_t *bar_ptr, uint32_t first_block, uint32_t lane)
    {
        base = packet;
        row = (uint32_t)__cvta_generic_to_shared(row_ptr);
        bars = (uint32_t)__cvta_generic_to_shared(bar_ptr);
        blocks = n / kBlockFrames;
        first = first_block;
        if (lane == 0) {
#pragma unroll
            for (uint32_t i = 0; i < kSlots; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" ::"r"(bars + 8u * i) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    __device__ __forceinline__ void request(uint32_t b) const
    {
        const uint32_t bar = bars + 8u * (b & (kSlots - 1u));
        if (b < blocks) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(G * 16u) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(row + (b & (kSlots - 1u)) * (G * 16u)), "l"(base + (size_t)b * (G * 16u)), "r"(G * 16u), "r"(bar) : "memory");
        } else {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
        }
    }
    __device__ __forceinline__ void acquire(uint32_t b) const
    {
        const uint32_t bar = bars + 8u * (b & (kSlots - 1u)), parity = ((b - first) / kSlots) & 1u;
        uint32_t done = 0;
        for (uint32_t spin = 0; !done; spin++) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar), "r"(parity) : "memory");
            if (spin > (1u << 24)) __trap();        // (an experiment must not hang the box)
        }
    }
    __device__ __forceinline__ void read(uint32_t b, uint32_t (&w)[WQ * QB]) const
    {
        const uint32_t src = row + (b & (kSlots - 1u)) * (G * 16u);
#pragma unroll
        for (uint32_t i = 0; i < G; i++) {
            const uint4 v = lds_u128(src + i * 16u);
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
        }
    }
};
#endif

// The lane's (mixed) samples of block b, in order: f(i, x) for i = 0 .. kBlockFrames - 1.  The quads of a block run
// through ONE copy of the caller's four-step body (the words of the later quads move down in registers): the bodies --
// predictor step plus Golomb step -- are long and these kernels feel instruction-cache pressure.
template <int DEPTH, bool STEREO, class R, class F>
__device__ __forceinline__ void ring_block_samples(const R &ring, uint32_t b, int32_t cl, int32_t cr, uint32_t sh_mix, F &&f)
{
    uint32_t w[R::WQ * R::QB];
    ring.read(b, w);
#pragma unroll 1
    for (uint32_t qi = 0; qi < R::QB; qi++) {
        uint32_t wq[R::WQ];
#pragma unroll
        for (uint32_t i = 0; i < R::WQ; i++) wq[i] = w[i];
        int32_t l[4], rr[4];
        unpack_quad<DEPTH, STEREO>(wq, l, rr);
#pragma unroll
        for (int i = 0; i < 4; i++) f(qi * 4u + (uint32_t)i, STEREO ? ((cl * l[i] + cr * rr[i]) >> sh_mix) : l[i]);
        if (R::QB > 1) {
#pragma unroll
            for (uint32_t i = 0; i + R::WQ < R::WQ * R::QB; i++) w[i] = w[i + R::WQ];
        }
    }
}

// A dense element's sample source for the search passes: the scalar MixSrc (warm-up samples, the odd frames at either
// end of a pass) plus the lane's word ring, through which predict_pass streams whole quads.
template <int DEPTH, bool STEREO>
struct DenseSrc : MixSrc<DEPTH, STEREO, true> {
    static constexpr bool kWide = false;
    static constexpr bool kDense = true;
    static constexpr int kDepth = DEPTH;
    static constexpr bool kStereo = STEREO;
    QuadRing<DEPTH, STEREO> q;
};

// pc_block(in, res, num, coefs, TAPS) streamed: sink(j, residual) is called for j = 0..max(num,TAPS+1)-1
// exactly as the reference writes pc1[j] (warm-up entries 1..TAPS are written regardless of num,
// codec/dp_enc.c:108-112).  Requires num <= src.valid (true for every caller), so the main loop
// needs no bounds checks; its loads run kAhead samples ahead of the arithmetic.
template <int TAPS, bool WRAP, class Src, class Sink>
__device__ __forceinline__ void predict_pass(const Src &src, uint32_t num, int32_t (&a)[TAPS], uint32_t chanshift, Sink &sink)
{
    int32_t hist[TAPS + 1];
    int32_t prev = src.get(0);
    sink(0u, prev);
    hist[TAPS] = prev;
#pragma unroll
    for (int j = 1; j <= TAPS; j++) {
        const int32_t x = src.get((uint32_t)j);
        sink((uint32_t)j, sext_bits(x - prev, chanshift));
        hist[TAPS - j] = x;
        prev = x;
    }
    if (num <= TAPS + 1) return;
    if constexpr (Src::kDense) {
        // frames up to the first block boundary after the warm-up one by one, whole blocks through the ring
        // (kQuadAhead blocks ahead of the arithmetic), then the odd frames after the last whole block of the pass
        constexpr int DEPTH = Src::kDepth;
        constexpr bool STEREO = Src::kStereo;
        using R = QuadRing<DEPTH, STEREO>;
        constexpr uint32_t b_first = (TAPS + 1 + R::kBlockFrames - 1) / R::kBlockFrames;
        const uint32_t nb = num / R::kBlockFrames;
        const R ring = src.q;
        if (nb > b_first) {
#pragma unroll
            for (uint32_t d = 0; d < kQuadAhead; d++) ring.request(b_first + d);
        }
        // (head and tail share ONE copy of the scalar loop -- two trips of the phase loop: the body is long, there are
        //  five instantiations of this pass in the search kernel, and the kernel feels instruction-cache pressure)
        uint32_t j = TAPS + 1;
        uint32_t stop = min(num, b_first * R::kBlockFrames);
#pragma unroll 1
        for (int phase = 0; phase < 2; phase++) {
#pragma unroll 1
            for (; j < stop; j++) sink(j, predict_enc_step<TAPS, WRAP>(src.get(j), hist, a, chanshift));
            stop = num;
            if (phase || nb <= b_first) continue;
            for (uint32_t b = b_first; b < nb; b++, j += R::kBlockFrames) {
                ring.request(b + kQuadAhead);
                cp_async_wait<kQuadAhead>();            // block b is in
                ring_block_samples<DEPTH, STEREO>(ring, b, src.cl, src.cr, src.sh_mix,
                                                  [&](uint32_t i, int32_t x) { sink(j + i, predict_enc_step<TAPS, WRAP>(x, hist, a, chanshift)); });
            }
            cp_async_wait<0>();
        }
        return;
    }
    if constexpr (Src::kWide) {
        // single frames up to the next 16-byte boundary of the PCM, then blocks of four, then the remaining frames.
        // Head and tail share ONE copy of the scalar loop (two trips of the phase loop): the body -- predictor step
        // plus Golomb step -- is long, and these kernels are sensitive to instruction-cache pressure.
        uint32_t j = TAPS + 1;
        const uint32_t a0 = (uint32_t)(reinterpret_cast<uintptr_t>(src.base) >> 2);
        uint32_t stop = min(num, j + ((0u - (a0 + j)) & 3u));
#pragma unroll 1
        for (int phase = 0; phase < 2; phase++) {
#pragma unroll 1
            for (; j < stop; j++) sink(j, predict_enc_step<TAPS, WRAP>(src.get(j), hist, a, chanshift));
            stop = num;
            if (phase) break;
            const uint32_t nblk = (num - j) >> 2;
            if (nblk) {
                // blocks of four frames stream through the lane's shared-memory ring, kPcmSlots - 1 blocks ahead;
                // unrolling by four also lets the history shift become register renaming
                const uint4 *p = reinterpret_cast<const uint4 *>(src.base + (size_t)j * 4u);
                const uint32_t slot_bytes = src.ring_stride;
#pragma unroll
                for (uint32_t d = 0; d + 1 < kPcmSlots; d++) {
                    cp_async_u128(src.ring + d * slot_bytes, p + min(d, nblk - 1u), d < nblk ? 16u : 0u);
                    cp_async_commit();
                }
                for (uint32_t b = 0; b < nblk; b++, j += 4) {
                    const uint32_t ahead = b + kPcmSlots - 1u;
                    cp_async_u128(src.ring + (ahead & (kPcmSlots - 1u)) * slot_bytes, p + min(ahead, nblk - 1u), ahead < nblk ? 16u : 0u);
                    cp_async_commit();
                    cp_async_wait<kPcmSlots - 1>();         // block b is in
                    const uint4 cur = lds_u128(src.ring + (b & (kPcmSlots - 1u)) * slot_bytes);
                    typename Src::Raw r;
                    r.w[0] = cur.x; sink(j, predict_enc_step<TAPS, WRAP>(src.mix(r), hist, a, chanshift));
                    r.w[0] = cur.y; sink(j + 1, predict_enc_step<TAPS, WRAP>(src.mix(r), hist, a, chanshift));
                    r.w[0] = cur.z; sink(j + 2, predict_enc_step<TAPS, WRAP>(src.mix(r), hist, a, chanshift));
                    r.w[0] = cur.w; sink(j + 3, predict_enc_step<TAPS, WRAP>(src.mix(r), hist, a, chanshift));
                }
                cp_async_wait<0>();
            }
        }
        return;
    }
    constexpr int D = Src::kAhead;
    const uint32_t last = num - 1;
    typename Src::Raw q[D];
#pragma unroll
    for (int d = 0; d < D; d++) q[d] = src.fetch(min((uint32_t)(TAPS + 1 + d), last));
    for (uint32_t j = TAPS + 1; j < num; j++) {
        const int32_t x = src.mix(q[0]);
#pragma unroll
        for (int d = 0; d + 1 < D; d++) q[d] = q[d + 1];
        q[D - 1] = src.fetch(min(j + D, last));
        const int32_t err = predict_enc_step<TAPS, WRAP>(x, hist, a, chanshift);
        sink(j, err);
    }
}

struct NullSink {
    __device__ __forceinline__ void operator()(uint32_t, int32_t) const {}
};

// Golomb bit-costing sink; optionally keeps the residuals (the stage-B "stale tail")
struct CostSink {
    AgEnc ag;
    uint32_t bit_size;
    int32_t *keep;      // may be null
    __device__ __forceinline__ void operator()(uint32_t j, int32_t err)
    {
        if (j < ag.count) {
            NoSink ns;
            ag_put<false>(ag, err, bit_size, ns);
            if (keep) keep[j] = err;
        }
    }
};

struct EmitSink {
    AgEnc ag;
    BitSink bits;
    uint32_t bit_size;
    __device__ __forceinline__ void operator()(uint32_t j, int32_t err)
    {
        if (j < ag.count) ag_put<true>(ag, err, bit_size, bits);
    }
};

__device__ __forceinline__ void init_coefs_row(int32_t *a, int n)
{
    // codec/dp_enc.c:49-60 with denshift 9
    for (int k = 0; k < n; k++) a[k] = 0;
    a[0] = (38 * 512) >> 4;
    a[1] = (-29 * 512) >> 4;
    a[2] = (-2 * 512) >> 4;
}

// Stages A and B of one frame for one chain lane (U and V of a pair on adjacent lanes): the mixRes search, the
// taps search and the escape estimate of EncodeStereo / EncodeMono.  c4 / c8 are the lane's coefficient rows.
struct SearchOut { uint32_t best_res, num_mine; int do_escape; };

template <int DEPTH, bool STEREO, bool WRAP, class Src>
__device__ __forceinline__ SearchOut search_stages(Src &src, const EncArgs &A, bool valid, bool is_v, uint32_t n,
                                                   uint32_t partial, uint32_t pair_mask, uint32_t chan_bits, uint32_t chanshift,
                                                   uint32_t *slab, int32_t (&c4)[4], int32_t (&c8)[8])
{
    constexpr uint32_t shift = DepthTraits<DEPTH>::kShift;
    uint32_t best_res = 0;
    uint32_t num_mine = 8;           // taps chosen for this lane's channel
    uint32_t metric_mine = 0;
    int do_escape = 0;

    const bool fast = STEREO && A.lay.fast_mode;   // EncodeMono has no fast variant
    if (!fast && valid) {
        if (STEREO) {
            // stage A: mixRes search, first n/8 samples, chained on row 7 (:353-379)
            const uint32_t na = n / 8;
            uint32_t min_bits = 1u << 31;
            for (int r = 0; r <= kMaxRes; r++) {
                src.set_mix(r, is_v);
                src.valid = na;
                CostSink cs;
                cs.ag.start(na);
                cs.bit_size = chan_bits;
                cs.keep = (r == kMaxRes) ? reinterpret_cast<int32_t *>(slab) : nullptr;
                predict_pass<8, WRAP>(src, na, c8, chanshift, cs);
                const uint32_t both = cs.ag.bits + __shfl_xor_sync(pair_mask, cs.ag.bits, 1);
                if (both < min_bits) { min_bits = both; best_res = (uint32_t)r; }
            }
            src.set_mix((int32_t)best_res, is_v);
            src.valid = n;
        }
        // stage B: taps search (:418-452 stereo, :881-905 mono)
        uint32_t best_metric = 1u << 31;
        num_mine = 4;
        const uint32_t nb = n / 32, nc = n / 8;
#pragma unroll 1
        for (uint32_t taps = 4; taps <= 8; taps += 4) {
            CostSink cs;
            cs.ag.start(nc);
            cs.bit_size = chan_bits;
            cs.keep = nullptr;
            NullSink null_sink;
            if (taps == 4) {
                for (int pass = 0; pass < 7; pass++) predict_pass<4, WRAP>(src, nb, c4, chanshift, null_sink);
                predict_pass<4, WRAP>(src, STEREO ? nb : nc, c4, chanshift, cs);
            } else {
                for (int pass = 0; pass < 7; pass++) predict_pass<8, WRAP>(src, nb, c8, chanshift, null_sink);
                predict_pass<8, WRAP>(src, STEREO ? nb : nc, c8, chanshift, cs);
            }
            if (STEREO) {
                // residuals [max(n/32, taps+1), n/8) are what the mixRes=4 trial left behind (F4)
                const int32_t *stale = reinterpret_cast<const int32_t *>(slab);
                NoSink ns;
                for (uint32_t j = max(nb, taps + 1); j < nc; j++) ag_put<false>(cs.ag, stale[j], chan_bits, ns);
            }
            const uint32_t metric = cs.ag.bits * 8 + 16 * taps;
            if (metric < best_metric) { best_metric = metric; num_mine = taps; }
        }
        metric_mine = best_metric;
        // escape estimate (:455-461 stereo, :909-915 mono)
        if (STEREO) {
            const uint32_t other = __shfl_xor_sync(pair_mask, metric_mine, 1);
            uint32_t min_bits = metric_mine + other + 64 + (partial ? 32u : 0u);
            if (shift) min_bits += n * shift * 2;
            const uint32_t escape_bits = n * DEPTH * 2 + (partial ? 32u : 0u) + 16;
            do_escape = (min_bits >= escape_bits);
        } else {
            uint32_t min_bits = metric_mine + 32 + (partial ? 32u : 0u);
            if (shift) min_bits += n * shift;
            const uint32_t escape_bits = n * DEPTH + (partial ? 32u : 0u) + 16;
            do_escape = (min_bits >= escape_bits);
        }
    }

    SearchOut o;
    o.best_res = best_res; o.num_mine = num_mine; o.do_escape = do_escape;
    return o;
}

// 128 chain lanes + one spare warp: the spare warp owns no chain; it only takes final-pass jobs so that
// the 4-tap and the 8-tap jobs can each start on a warp boundary and no warp ever runs both loops.
#ifndef ALAC_CHAIN_THREADS
#define ALAC_CHAIN_THREADS 96
#endif
constexpr int kChainThreads = ALAC_CHAIN_THREADS;
#ifndef ALAC_SPARE_WARPS
#define ALAC_SPARE_WARPS 1
#endif
constexpr int kSearchThreads = kChainThreads + 32 * ALAC_SPARE_WARPS;

// What the final pass (stage C) of one channel needs; lets any lane of the CTA run it.
struct FinalJob {
    const uint8_t *base;    // sample-frame 0 of the packet, first channel of the element
    uint32_t *slab;         // Golomb stream destination
    int32_t coef[8];        // the selected coefficient row (in: after the search, out: after the pass)
    uint32_t *bits_out;     // where the pass reports its Golomb bit count (ElemRec::bits_u / bits_v)
    uint32_t n;             // samples in the frame
    uint32_t flags;         // bit 0: V channel, bits 1..3: mixRes
};

template <int DEPTH, bool STEREO, bool PACKED, bool WRAP>
__global__ void __launch_bounds__(kSearchThreads, 864 / kSearchThreads)
enc_search_kernel(EncArgs A, uint32_t elems_of_kind, uint32_t kind_elem0 /* bitmask of element slots of this kind */)
{
    // Stages A and B keep the U and V chains of a pair on adjacent lanes (they trade bit counts by
    // shuffle and always run the same tap count).  The final pass runs 4 OR 8 taps per channel, so
    // before it the CTA's chains are regrouped by tap count through shared memory: lanes [0, n4) run
    // the 4-tap pass, lanes [n4, n4 + n8) the 8-tap pass, and each warp stays uniform.
    __shared__ FinalJob s_job[kSearchThreads];
    __shared__ uint32_t s_cnt[kSearchThreads / 32][2];
    __shared__ PcmRing<kSearchThreads> s_pcm;
    __shared__ uint32_t s_pn_max;

    constexpr uint32_t kLanesPerJob = STEREO ? 2 : 1;
    const uint32_t tid = blockIdx.x * kChainThreads + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint32_t job = tid / kLanesPerJob;
    const bool is_v = STEREO && (tid & 1u);
    const uint32_t total_jobs = A.num_segments * elems_of_kind;
    // lanes of a pair are both in or both out; the spare warp(s) own no chain
    const bool in_range = threadIdx.x < kChainThreads && job < total_jobs;
    const uint32_t pair_mask = STEREO ? (3u << (lane & ~1u)) : 0u;

    uint32_t slot = 0, seg = A.seg_base;
    if (in_range) {
        seg = A.seg_base + job / elems_of_kind;
        uint32_t which = job % elems_of_kind;
        // slot = index of the which-th element of this kind inside the packet
        for (uint32_t s = 0, m = kind_elem0; s < 8; s++, m >>= 1) {
            if (m & 1u) { if (which == 0) { slot = s; break; } which--; }
        }
    }
    const uint32_t chan = A.lay.elem_chan[slot];
    const uint32_t chain = A.lay.elem_chain[slot] + (is_v ? 1u : 0u);
    constexpr uint32_t bps = DepthTraits<DEPTH>::kBytes;
    constexpr uint32_t shift = DepthTraits<DEPTH>::kShift;
    const uint32_t stride = A.lay.channels * bps;
    const uint32_t chan_bits = DEPTH - shift + (STEREO ? 1u : 0u);                 // :334 / :857
    const uint32_t chanshift = 32u - chan_bits;

    // coefficient rows 3 (4 taps) and 7 (8 taps) of this channel (codec/ALACEncoder.h:89-90)
    int32_t c4[4], c8[8];
    const uint32_t seg_info = in_range ? A.seg_stream[seg] : 0u;
    const uint32_t stream = seg_info & 0x3fffffffu;
    int16_t *st = (A.state && in_range) ? A.state + ((size_t)stream * 8 + chan) * 32 + (is_v ? 16 : 0) : nullptr;
    if (st && (seg_info & 0x80000000u)) {
        for (int k = 0; k < 4; k++) c4[k] = st[k];
        for (int k = 0; k < 8; k++) c8[k] = st[8 + k];
    } else {
        init_coefs_row(c4, 4);
        init_coefs_row(c8, 8);
    }

    const uint32_t p0 = in_range ? A.seg_first[seg] : 0u;
    const uint32_t pn = in_range ? A.seg_count[seg] : 0u;
    if (threadIdx.x == 0) s_pn_max = 0;
    __syncthreads();
    atomicMax(&s_pn_max, pn);
    __syncthreads();
    const uint32_t pn_max = s_pn_max;               // CTA-uniform trip count (barriers inside the loop)

    for (uint32_t pi = 0; pi < pn_max; pi++) {
        const bool valid = pi < pn;
        const uint32_t pkt = valid ? p0 + pi : A.pkt_base;
        const uint32_t n = valid ? A.pkt_samples[pkt] : 0u;
        const uint8_t *base = A.pcm + (A.pkt_frame[pkt] * A.lay.channels + chan) * bps;
        uint32_t *slab = A.scratch + ((size_t)(pkt - A.pkt_base) * A.lay.chains_per_packet + chain) * A.cap_words;
        const uint32_t partial = (n != A.lay.frame_size);

        MixSrc<DEPTH, STEREO, PACKED> src;
        src.base = base; src.stride = stride; src.valid = n;
        src.ring = pcm_ring_addr(s_pcm); src.ring_stride = (uint32_t)sizeof(uint4) * kSearchThreads;
        src.set_mix(0, is_v);

        const SearchOut so = search_stages<DEPTH, STEREO, WRAP>(src, A, valid, is_v, n, partial, pair_mask, chan_bits, chanshift, slab, c4, c8);
        const uint32_t best_res = so.best_res, num_mine = so.num_mine;
        const int do_escape = so.do_escape;

        // header coefficients are the post-search, pre-final-pass values (:479-485)
        ElemRec *rec = A.recs + (size_t)(pkt - A.pkt_base) * A.lay.elems_per_packet + slot;
        const uint32_t other_num = STEREO ? __shfl_xor_sync(0xffffffffu, num_mine, 1) : 0u;
        if (valid) {
            int16_t *hc = is_v ? rec->coef_v : rec->coef_u;
            if (num_mine == 4) { for (int k = 0; k < 4; k++) hc[k] = (int16_t)c4[k]; for (int k = 4; k < 8; k++) hc[k] = 0; }
            else { for (int k = 0; k < 8; k++) hc[k] = (int16_t)c8[k]; }
            // the post-check and the element size are finished by enc_size_kernel once both bit counts exist
            if (is_v) {
                rec->bits_v = 0;
            } else {
                rec->escape = (uint8_t)do_escape;
                rec->mix_res = (uint8_t)best_res;
                rec->num_u = (uint8_t)num_mine;
                rec->num_v = (uint8_t)other_num;
                rec->bits_u = 0;
                if (!STEREO) rec->bits_v = 0;
            }
        }

        // ---- stage C: final predictor + Golomb pass over the whole frame (:507-531, :941-945),
        //      regrouped by tap count across the CTA
        const uint32_t key = (valid && !do_escape) ? (num_mine == 8 ? 1u : 0u) : 2u;
        const uint32_t b4 = __ballot_sync(0xffffffffu, key == 0), b8 = __ballot_sync(0xffffffffu, key == 1);
        if (lane == 0) { s_cnt[wid][0] = __popc(b4); s_cnt[wid][1] = __popc(b8); }
        __syncthreads();
        uint32_t n4 = 0, n8 = 0, before4 = 0, before8 = 0;
#pragma unroll
        for (uint32_t w = 0; w < kSearchThreads / 32; w++) {
            if (w < wid) { before4 += s_cnt[w][0]; before8 += s_cnt[w][1]; }
            n4 += s_cnt[w][0];
            n8 += s_cnt[w][1];
        }
        const uint32_t lt = (1u << lane) - 1u;
        // 8-tap jobs start on the next warp boundary after the 4-tap jobs (fits thanks to the spare warp)
        const uint32_t first8 = ALAC_SPARE_WARPS ? ((n4 + 31u) & ~31u) : n4;
        const uint32_t my_slot = key == 0 ? before4 + __popc(b4 & lt) : first8 + before8 + __popc(b8 & lt);
        if (key < 2) {
            FinalJob &J = s_job[my_slot];
            J.base = base;
            J.slab = slab;
            J.bits_out = is_v ? &rec->bits_v : &rec->bits_u;
            J.n = n;
            J.flags = (is_v ? 1u : 0u) | (best_res << 1);
            if (key == 0) { for (int k = 0; k < 4; k++) J.coef[k] = c4[k]; }
            else { for (int k = 0; k < 8; k++) J.coef[k] = c8[k]; }
        }
        __syncthreads();
        if (threadIdx.x < n4 || (threadIdx.x >= first8 && threadIdx.x < first8 + n8)) {
            FinalJob &J = s_job[threadIdx.x];
            MixSrc<DEPTH, STEREO, PACKED> fs;
            fs.base = J.base; fs.stride = stride;
            fs.ring = pcm_ring_addr(s_pcm); fs.ring_stride = (uint32_t)sizeof(uint4) * kSearchThreads;
            fs.set_mix((int32_t)(J.flags >> 1), (J.flags & 1u) != 0);
            fs.valid = J.n;
            EmitSink es;
            es.ag.start(J.n);
            es.bit_size = chan_bits;
            es.bits.start(J.slab, A.cap_words);
            if (threadIdx.x < n4) {
                int32_t a[4];
                for (int k = 0; k < 4; k++) a[k] = J.coef[k];
                predict_pass<4, WRAP>(fs, J.n, a, chanshift, es);
                for (int k = 0; k < 4; k++) J.coef[k] = a[k];
            } else {
                int32_t a[8];
                for (int k = 0; k < 8; k++) a[k] = J.coef[k];
                predict_pass<8, WRAP>(fs, J.n, a, chanshift, es);
                for (int k = 0; k < 8; k++) J.coef[k] = a[k];
            }
            es.bits.finish();
            *J.bits_out = es.ag.bits;
        }
        // the adapted coefficients only matter if another frame of the segment follows or the state is exported
        if (pi + 1 < pn_max || A.state != nullptr) {
            __syncthreads();
            if (key < 2) {
                const FinalJob &J = s_job[my_slot];
                if (key == 0) { for (int k = 0; k < 4; k++) c4[k] = J.coef[k]; }
                else { for (int k = 0; k < 8; k++) c8[k] = J.coef[k]; }
            }
        }

    }

    if (st && (seg_info & 0x40000000u)) {
        for (int k = 0; k < 4; k++) st[k] = (int16_t)c4[k];
        for (int k = 4; k < 8; k++) st[k] = 0;
        for (int k = 0; k < 8; k++) st[8 + k] = (int16_t)c8[k];
    }
}

// ---- split form of the search kernel (every segment is one frame: frames_per_segment = 1, no state hand-off) -------
// Nothing chains from one frame to the next, so the final pass need not stay on the lane (or even in the CTA) that
// did the search.  enc_search_split_kernel runs stages A and B with one-warp CTAs (the grid then spreads over the
// 148 SMs to within one warp) and files each channel's final-pass job under its tap count in a global list;
// enc_final_kernel runs stage C with every warp full and uniform -- no in-CTA regrouping, no barriers, no spare warp.
struct JobLists {
    FinalJob *jobs;         // [2][max_jobs]: 4-tap jobs, then 8-tap jobs
    uint32_t *counts;       // [2]
    uint32_t max_jobs;
};

// DENSE: a dense element (see DenseElem) -- the passes stream their PCM through the lane's word ring instead of
// per-sample loads (20/24/32-bit and mono streams; 16-bit packed stereo has its own 16-byte ring, MixSrc::kWide).
#ifndef ALAC_SEARCH_MINB
#define ALAC_SEARCH_MINB 20
#endif
#ifndef ALAC_FINAL1_MINB
#define ALAC_FINAL1_MINB 24
#endif
template <int DEPTH, bool STEREO, bool PACKED, bool WRAP, bool DENSE = false>
__global__ void __launch_bounds__(32, ALAC_SEARCH_MINB)
enc_search_split_kernel(EncArgs A, uint32_t elems_of_kind, uint32_t kind_elem0, JobLists Q)
{
    constexpr uint32_t kLanesPerJob = STEREO ? 2 : 1;
    const uint32_t tid = blockIdx.x * 32u + threadIdx.x;
    const uint32_t lane = threadIdx.x;
    const uint32_t job = tid / kLanesPerJob;
    const bool is_v = STEREO && (tid & 1u);
    const uint32_t total_jobs = A.num_segments * elems_of_kind;
    const bool valid = job < total_jobs;              // lanes of a pair are both in or both out
    const uint32_t pair_mask = STEREO ? (3u << (lane & ~1u)) : 0u;

    uint32_t slot = 0, seg = A.seg_base;
    if (valid) {
        seg = A.seg_base + job / elems_of_kind;
        uint32_t which = job % elems_of_kind;
        for (uint32_t s = 0, m = kind_elem0; s < 8; s++, m >>= 1) {
            if (m & 1u) { if (which == 0) { slot = s; break; } which--; }
        }
    }
    const uint32_t chan = A.lay.elem_chan[slot];
    const uint32_t chain = A.lay.elem_chain[slot] + (is_v ? 1u : 0u);
    constexpr uint32_t bps = DepthTraits<DEPTH>::kBytes;
    constexpr uint32_t shift = DepthTraits<DEPTH>::kShift;
    const uint32_t stride = A.lay.channels * bps;
    const uint32_t chan_bits = DEPTH - shift + (STEREO ? 1u : 0u);                 // :334 / :857
    const uint32_t chanshift = 32u - chan_bits;

    int32_t c4[4], c8[8];
    init_coefs_row(c4, 4);
    init_coefs_row(c8, 8);

    const uint32_t pkt = valid ? A.seg_first[seg] : A.pkt_base;
    const uint32_t n = valid ? A.pkt_samples[pkt] : 0u;
    const uint8_t *base = A.pcm + (A.pkt_frame[pkt] * A.lay.channels + chan) * bps;
    uint32_t *slab = A.scratch + ((size_t)(pkt - A.pkt_base) * A.lay.chains_per_packet + chain) * A.cap_words;
    const uint32_t partial = (n != A.lay.frame_size);

    __shared__ uint4 s_ring[DENSE ? QuadRing<DEPTH, STEREO>::kRows : kPcmSlots][32];     // DENSE: block ring; else the 16-byte PCM ring of MixSrc::kWide
    typename std::conditional<DENSE, DenseSrc<DEPTH, STEREO>, MixSrc<DEPTH, STEREO, PACKED>>::type src;
    src.base = base; src.stride = stride; src.valid = n;
    src.ring = (uint32_t)__cvta_generic_to_shared(&s_ring[0][0]) + threadIdx.x * 16u; src.ring_stride = (uint32_t)sizeof(uint4) * 32u;
    if constexpr (DENSE) src.q.start(valid ? base : reinterpret_cast<const uint8_t *>(&s_ring[0][0]), valid ? n : 0u, &s_ring[0][lane]);
    src.set_mix(0, is_v);
    const SearchOut so = search_stages<DEPTH, STEREO, WRAP>(src, A, valid, is_v, n, partial, pair_mask, chan_bits, chanshift, slab, c4, c8);
    const uint32_t best_res = so.best_res, num_mine = so.num_mine;
    const int do_escape = so.do_escape;

    // header coefficients are the post-search, pre-final-pass values (:479-485)
    ElemRec *rec = A.recs + (size_t)(pkt - A.pkt_base) * A.lay.elems_per_packet + slot;
    const uint32_t other_num = STEREO ? __shfl_xor_sync(0xffffffffu, num_mine, 1) : 0u;
    if (valid) {
        int16_t *hc = is_v ? rec->coef_v : rec->coef_u;
        if (num_mine == 4) { for (int k = 0; k < 4; k++) hc[k] = (int16_t)c4[k]; for (int k = 4; k < 8; k++) hc[k] = 0; }
        else { for (int k = 0; k < 8; k++) hc[k] = (int16_t)c8[k]; }
        if (is_v) {
            rec->bits_v = 0;
        } else {
            rec->escape = (uint8_t)do_escape;
            rec->mix_res = (uint8_t)best_res;
            rec->num_u = (uint8_t)num_mine;
            rec->num_v = (uint8_t)other_num;
            rec->bits_u = 0;
            if (!STEREO) rec->bits_v = 0;
        }
    }

    // file the final-pass job under its tap count (one atomic per warp and list)
    const uint32_t key = (valid && !do_escape) ? (num_mine == 8 ? 1u : 0u) : 2u;
    const uint32_t b4 = __ballot_sync(0xffffffffu, key == 0), b8 = __ballot_sync(0xffffffffu, key == 1);
    uint32_t first4 = 0, first8 = 0;
    if (lane == 0) {
        if (b4) first4 = atomicAdd(&Q.counts[0], (uint32_t)__popc(b4));
        if (b8) first8 = atomicAdd(&Q.counts[1], (uint32_t)__popc(b8));
    }
    first4 = __shfl_sync(0xffffffffu, first4, 0);
    first8 = __shfl_sync(0xffffffffu, first8, 0);
    if (key < 2) {
        const uint32_t lt = (1u << lane) - 1u;
        FinalJob J;
        J.base = base;
        J.slab = slab;
        J.bits_out = is_v ? &rec->bits_v : &rec->bits_u;
        J.n = n;
        J.flags = (is_v ? 1u : 0u) | (best_res << 1);
#pragma unroll
        for (int k = 0; k < 8; k++) J.coef[k] = (key == 0) ? (k < 4 ? c4[k] : 0) : c8[k];
        Q.jobs[(size_t)key * Q.max_jobs + (key == 0 ? first4 + __popc(b4 & lt) : first8 + __popc(b8 & lt))] = J;
    }
}

// stage C: final predictor + Golomb pass over the whole frame (:507-531, :941-945), one lane per listed job.
// The first half of the grid takes the 8-tap list (the longer jobs start first), the second half the 4-tap list.
template <int DEPTH, bool STEREO, bool PACKED, bool WRAP>
__global__ void __launch_bounds__(32, 24)
enc_final_kernel(EncArgs A, JobLists Q, uint32_t ctas_per_list)
{
    const uint32_t list = blockIdx.x < ctas_per_list ? 1u : 0u;
    const uint32_t idx = (blockIdx.x - (list ? 0u : ctas_per_list)) * 32u + threadIdx.x;
    if (idx >= Q.counts[list]) return;
    const FinalJob J = Q.jobs[(size_t)list * Q.max_jobs + idx];
    constexpr uint32_t bps = DepthTraits<DEPTH>::kBytes;
    constexpr uint32_t shift = DepthTraits<DEPTH>::kShift;
    const uint32_t chan_bits = DEPTH - shift + (STEREO ? 1u : 0u);
    const uint32_t chanshift = 32u - chan_bits;
    __shared__ PcmRing<32> s_pcm;
    MixSrc<DEPTH, STEREO, PACKED> fs;
    fs.base = J.base; fs.stride = A.lay.channels * bps;
    fs.ring = pcm_ring_addr(s_pcm); fs.ring_stride = (uint32_t)sizeof(uint4) * 32u;
    fs.set_mix((int32_t)(J.flags >> 1), (J.flags & 1u) != 0);
    fs.valid = J.n;
    EmitSink es;
    es.ag.start(J.n);
    es.bit_size = chan_bits;
    es.bits.start(J.slab, A.cap_words);
    if (list == 0) {
        int32_t a[4];
        for (int k = 0; k < 4; k++) a[k] = J.coef[k];
        predict_pass<4, WRAP>(fs, J.n, a, chanshift, es);
    } else {
        int32_t a[8];
        for (int k = 0; k < 8; k++) a[k] = J.coef[k];
        predict_pass<8, WRAP>(fs, J.n, a, chanshift, es);
    }
    es.bits.finish();
    *J.bits_out = es.ag.bits;
}

// ---- two-warp form of the final pass: predictor warp -> residual tiles -> Golomb warp ------------------------------
// One lane per job as in enc_final_kernel, but the two serial chains of a job -- the sign-LMS predictor and the
// adaptive Golomb coder -- run on two warps of a 64-thread CTA, lane = job in both, handing tiles of 32 residuals per
// lane through shared memory (FULL / EMPTY named barriers per buffer, as dec_fused_kernel does on the decode side).
// The per-sample critical path becomes max(predictor, coder) instead of their sum, twice as many warps are in
// flight, and each warp's loop body is about half as long.
//
// "Dense" elements only: a mono or stereo stream whose element is the whole sample-frame (no other channels in
// between) and whose packets start on 4-byte boundaries.  Then a lane's PCM is one contiguous run of 32-bit words and
// streams through a word ring in shared memory -- quad q of lane l in rows [(q mod 4) * WQ, + WQ) of column l, so every warp access
// is conflict-free -- filled by 4-byte cp.async kQuadAhead quads (4 sample-frames each) ahead of the arithmetic.
// 24-bit frames (6 bytes) unpack from the ring with one PRMT per sample (bytes o+1, o+2 and the sign of o+2: the
// predictor input is sample >> 8); the per-sample byte loads of the generic path are gone.
// SPLIT = true: the two-warp form (64-thread CTAs).  SPLIT = false: the same dense PCM path with predictor and coder
// on ONE warp (32-thread CTAs, no residual tiles, no barriers): once a launch holds several waves of jobs the GPU is
// throughput-bound, the coder warp's idle half only costs occupancy, and the one-warp form is the faster one
// (10-hour 24/96 corpus: 44 ms against 52 ms; 1-hour 16/44.1: 2.9 ms against 2.6 ms).  The engine picks by job count.
template <int DEPTH, bool STEREO, bool WRAP, bool SPLIT>
__global__ void __launch_bounds__(SPLIT ? 64 : 32, SPLIT ? 12 : ALAC_FINAL1_MINB)
enc_final2_kernel(EncArgs A, JobLists Q, uint32_t ctas_per_list)
{
    __shared__ uint4 s_ring[QuadRing<DEPTH, STEREO>::kRows][32];
    __shared__ int32_t s_res[SPLIT ? 2 : 1][SPLIT ? kEncTileRows : 1][32];

    const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t list = blockIdx.x < ctas_per_list ? 1u : 0u;
    const uint32_t idx = (blockIdx.x - (list ? 0u : ctas_per_list)) * 32u + lane;
    const uint32_t count = Q.counts[list];
    if ((blockIdx.x - (list ? 0u : ctas_per_list)) * 32u >= count) return;     // CTA-uniform
    const bool have = idx < count;
    FinalJob J;
    if (have) J = Q.jobs[(size_t)list * Q.max_jobs + idx];
    else { J.base = nullptr; J.slab = nullptr; J.bits_out = nullptr; J.n = 0; J.flags = 0; }
    constexpr uint32_t shift = DepthTraits<DEPTH>::kShift;
    const uint32_t chan_bits = DEPTH - shift + (STEREO ? 1u : 0u);
    const uint32_t chanshift = 32u - chan_bits;
    const uint32_t n = J.n;
    const uint32_t n_max = __reduce_max_sync(0xffffffffu, n);
    const uint32_t tiles = (n_max + kEncTileRows - 1u) / kEncTileRows;

    EmitSink es;
    es.ag.start(n);
    es.bit_size = chan_bits;
    es.bits.start(J.slab, A.cap_words);
    if (SPLIT && w == 1) {
        // ================= Golomb warp =================
        for (uint32_t b = 0; b < min(tiles, 2u); b++) named_arrive<BAR_EMPTY0>(b != 0);     // both buffers start empty
        for (uint32_t t = 0; t < tiles; t++) {
            const uint32_t b = t & 1u;
            named_sync<BAR_FULL0>(b != 0);
            const int32_t *col = &s_res[SPLIT ? b : 0][0][lane];
            const uint32_t j0 = t * kEncTileRows;
            if (j0 + kEncTileRows <= n) {
#pragma unroll 2
                for (uint32_t r = 0; r < kEncTileRows; r++) ag_put<true>(es.ag, col[r * 32u], chan_bits, es.bits);
            } else {
                for (uint32_t r = 0; r < kEncTileRows && j0 + r < n; r++) ag_put<true>(es.ag, col[r * 32u], chan_bits, es.bits);
            }
            __syncwarp();
            if (t + 2 < tiles) named_arrive<BAR_EMPTY0>(b != 0);
        }
        if (have) {
            es.bits.finish();
            *J.bits_out = es.ag.bits;
        }
        return;
    }

    // ================= predictor warp (SPLIT) / the whole pass (one-warp form) =================
    MixSrc<DEPTH, STEREO, true> src;        // the scalar path for the warm-up samples and the odd frames at either end
    src.base = J.base; src.stride = DenseElem<DEPTH, STEREO>::kFrameBytes;
    src.set_mix((int32_t)(J.flags >> 1), (J.flags & 1u) != 0);
    src.valid = n;
#ifdef ALAC_RING_BULK
    using R = BulkRing<DEPTH, STEREO>;
    __shared__ __align__(16) uint8_t s_bulk[32][R::kRowBytes];
    __shared__ __align__(8) uint64_t s_bar[R::kSlots];
    R ring;
    ring.start(have ? J.base : reinterpret_cast<const uint8_t *>(s_ring), n, &s_bulk[lane][0], &s_bar[0], ((list ? 8u : 4u) + 1u + R::kBlockFrames - 1u) / R::kBlockFrames, lane);
#else
    QuadRing<DEPTH, STEREO> ring;
    ring.start(have ? J.base : reinterpret_cast<const uint8_t *>(s_ring), n, &s_ring[0][lane]);
    using R = QuadRing<DEPTH, STEREO>;
#endif
    static_assert(kEncTileRows % R::kBlockFrames == 0, "a tile is a whole number of blocks");
    const uint32_t nblk = ring.blocks;      // whole blocks of the packet
    auto run = [&](auto taps_tag) {
        constexpr int TAPS = decltype(taps_tag)::value;
        int32_t a[TAPS], hist[TAPS + 1];
#pragma unroll
        for (int k = 0; k < TAPS; k++) a[k] = J.coef[k];
#pragma unroll
        for (int k = 0; k <= TAPS; k++) hist[k] = 0;
        // blocks b_first .. b_first + kQuadAhead - 1 are requested up front; the warm-up samples and the frames up to
        // the first block boundary after them go through the scalar path
        constexpr uint32_t b_first = (TAPS + 1 + R::kBlockFrames - 1) / R::kBlockFrames;
#pragma unroll
        for (uint32_t d = 0; d < kQuadAhead; d++) ring.request(b_first + d);
        int32_t *col = &s_res[0][0][lane];
        // residual of sample-frame j0 + r: to the tile (two-warp form) or straight into the coder
        auto put = [&](uint32_t r, uint32_t j, int32_t err) {
            if (SPLIT) col[r * 32u] = err;
            else if (j < n) ag_put<true>(es.ag, err, chan_bits, es.bits);
        };
        int32_t prev = 0;
        for (uint32_t t = 0; t < tiles; t++) {
            const uint32_t b = t & 1u;
            if (SPLIT) {
                named_sync<BAR_EMPTY0>(b != 0);
                col = &s_res[SPLIT ? b : 0][0][lane];
            }
            const uint32_t j0 = t * kEncTileRows;
            uint32_t r = 0;
            if (t == 0) {
                // warm-up: pc[0] = x[0], pc[1..TAPS] = first differences, written regardless of n (dp_enc.c:108-112)
                prev = src.get(0);
                put(0, 0, prev);
                hist[TAPS] = prev;
#pragma unroll
                for (int j = 1; j <= TAPS; j++) {
                    const int32_t x = src.get((uint32_t)j);
                    put((uint32_t)j, (uint32_t)j, sext_bits(x - prev, chanshift));
                    hist[TAPS - j] = x;
                    prev = x;
                }
                r = TAPS + 1;
            }
            // whole blocks of this tile through the ring; the frames between the warm-up and the first block boundary and
            // the odd frames after the last whole block one by one (ONE copy of the scalar step: instruction cache)
#pragma unroll 1
            while (r < kEncTileRows) {
                const uint32_t j = j0 + r, blk = j / R::kBlockFrames;
                if (j >= b_first * R::kBlockFrames && blk < nblk) {     // (j is on a block boundary here)
                    ring.request(blk + kQuadAhead);
#ifdef ALAC_RING_BULK
                    ring.acquire(blk);
#else
                    cp_async_wait<kQuadAhead>();            // block blk is in
#endif
                    ring_block_samples<DEPTH, STEREO>(ring, blk, src.cl, src.cr, src.sh_mix, [&](uint32_t i, int32_t x) {
                        put(r + i, j + i, predict_enc_step<TAPS, WRAP>(x, hist, a, chanshift));
                    });
                    r += R::kBlockFrames;
                } else {
#ifdef ALAC_RING_BULK
                    // (a lane past its last whole block still arrives once per block index: the barrier counts all 32 lanes)
                    if (j >= b_first * R::kBlockFrames && j % R::kBlockFrames == 0) ring.request(blk + kQuadAhead);
#endif
                    if (j < n) put(r, j, predict_enc_step<TAPS, WRAP>(src.get(j), hist, a, chanshift));
                    r++;
                }
            }
            if (SPLIT) {
                __syncwarp();
                named_arrive<BAR_FULL0>(b != 0);
            }
        }
        cp_async_wait<0>();
    };
    if (list == 0) run(std::integral_constant<int, 4>{});
    else run(std::integral_constant<int, 8>{});
    if (!SPLIT && have) {
        es.bits.finish();
        *J.bits_out = es.ag.bits;
    }
}

// ---- packet sizes ------------------------------------------------------------------------------------
// Finishes each element record (post-check of codec/ALACEncoder.cu:537-543 / :952-958, fast mode
// :703-725: a compressed element that is not smaller than the escape form is sent as escape) and
// sums the element bits of a packet.
static __global__ void enc_size_kernel(ElemRec *recs, EncLayout lay, uint32_t depth, const uint32_t *pkt_samples,
                                uint32_t num_packets, uint32_t *sizes, unsigned long long *escapes)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= num_packets) return;
    const uint32_t n = pkt_samples[p];
    const uint32_t partial = (n != lay.frame_size) ? 32u : 0u;
    const uint32_t shift = depth == 32 ? 16u : depth == 24 ? 8u : 0u;
    uint32_t bits = 3;      // ID_END, codec/ALACEncoder.cu:1036
    uint32_t esc = 0;
    for (uint32_t e = 0; e < lay.elems_per_packet; e++) {
        ElemRec &r = recs[(size_t)p * lay.elems_per_packet + e];
        const bool stereo = (lay.elem_tag[e] == ID_CPE);
        const uint32_t nch = stereo ? 2u : 1u;
        const uint32_t escape_bits = n * depth * nch + partial + 16;
        const uint32_t body_bits = 12 + 4 + partial + 16 + (16 + 16 * r.num_u) + (stereo ? 16 + 16 * r.num_v : 0u) +
                                   n * shift * nch + r.bits_u + (stereo ? r.bits_v : 0u);
        uint32_t do_escape = r.escape;
        if (!do_escape) {
            if (stereo && lay.fast_mode) {
                const uint32_t min_bits = (r.bits_u + r.num_u * 16) + (r.bits_v + r.num_v * 16) + 64 + partial + n * shift * 2;
                if (min_bits >= escape_bits) do_escape = 2;
            }
            if (!do_escape && body_bits >= escape_bits) do_escape = 2;
        }
        r.escape = (uint8_t)do_escape;
        r.elem_bits = 7 + (do_escape ? (12 + 4 + partial + n * depth * nch) : body_bits);
        bits += r.elem_bits;
        esc += do_escape ? 1u : 0u;
    }
    sizes[p] = (bits + 7) >> 3;     // byte-align, :1039
    if (esc) atomicAdd(escapes, (unsigned long long)esc);
}

// ---- exclusive scan (two launches, any length) ----------------------------------------------------------------
// offsets[i] = sum_{j<i} sizes[j]; offsets[n] = total.  Tiles of 4096 elements (four consecutive elements per thread of a
// 1024-thread block): scan_tile_sums_kernel reduces every tile, scan_u32_to_u64_kernel sums the tile totals in front
// of its tile (a few hundred values even for the 843,750 packets of a 10-hour stream), scans its own tile and writes it.
constexpr uint32_t kScanTile = 4096;
__device__ __forceinline__ uint64_t block_sum_u64(uint64_t x, uint64_t *warp_sums)
{
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
    __syncthreads();                // warp_sums may still be read from a previous use
    if (lane == 0) warp_sums[wid] = x;
    __syncthreads();
    uint64_t s = warp_sums[lane];   // 32 warps
#pragma unroll
    for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    return s;                       // every thread holds the block total
}

static __global__ void __launch_bounds__(1024) scan_tile_sums_kernel(const uint32_t *in, uint64_t n, uint64_t *tile_sums)
{
    __shared__ uint64_t warp_sums[32];
    const uint64_t i0 = (uint64_t)blockIdx.x * kScanTile + 4ull * threadIdx.x;
    uint64_t x = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) x += i0 + k < n ? in[i0 + k] : 0u;
    const uint64_t total = block_sum_u64(x, warp_sums);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// chain_base: continue from out[0], which the previous chunk's scan left as its grand total (tile 0 rewrites out[0]
// with that same value, so tiles reading it concurrently see one value either way).
static __global__ void __launch_bounds__(1024) scan_u32_to_u64_kernel(const uint32_t *in, uint64_t *out, uint64_t n, const uint64_t *tile_sums,
                                                               uint32_t *max_out, int chain_base, uint64_t *host_total = nullptr)
{
    __shared__ uint64_t warp_sums[32];
    __shared__ uint32_t warp_max[32];
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    // tile totals in front of this tile
    uint64_t before = 0;
    for (uint32_t t = threadIdx.x; t < blockIdx.x; t += 1024u) before += tile_sums[t];
    uint64_t carry = block_sum_u64(before, warp_sums);
    if (chain_base) carry += *reinterpret_cast<volatile const uint64_t *>(out);
    const uint64_t i0 = (uint64_t)blockIdx.x * kScanTile + 4ull * threadIdx.x;
    uint32_t v[4];
    uint32_t my_max = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { v[k] = i0 + k < n ? in[i0 + k] : 0u; my_max = max(my_max, v[k]); }
    uint64_t x = (uint64_t)v[0] + v[1] + v[2] + v[3];
    const uint64_t mine = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= (uint32_t)d) x += y;
    }
    __syncthreads();
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint64_t s = warp_sums[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t y = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= (uint32_t)d) s += y;
        }
        warp_sums[lane] = s;
    }
    __syncthreads();
    const uint64_t incl = x + (wid ? warp_sums[wid - 1] : 0) + carry;
    uint64_t run = incl - mine;         // exclusive prefix of this thread's first element
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (i0 + k < n) out[i0 + k] = run;
        run += v[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 1023) {
        out[n] = incl;
        // pinned host word (UVA): the host reads it after the chunk's completion event, no copy-engine op needed
        if (host_total) { *host_total = incl; __threadfence_system(); }
    }
    if (max_out) {
        for (int d = 16; d; d >>= 1) my_max = max(my_max, __shfl_xor_sync(0xffffffffu, my_max, d));
        if (lane == 0) warp_max[wid] = my_max;
        __syncthreads();
        if (wid == 0) {
            uint32_t m = warp_max[lane];
            for (int d = 16; d; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
            if (lane == 0) atomicMax(max_out, m);
        }
    }
}

// ---- cross-GPU packet-offset exchange -----------------------------------------------------------------
// A 1 KB block in the destination GPU's memory.  Rank r publishes the byte total of its packet block, then waits -- on
// the device -- for the totals of the ranks in front of it: their sum is where its block starts.  Slots are tagged with
// the call's epoch and alternate by its parity; the home rank ends a call only when every rank has reported `done`, so no
// rank can be more than one call ahead of another and two slots suffice.
struct Exchange {
    unsigned long long total[2][16];
    uint32_t ready[2][16];
    uint32_t done[2][16];
    uint32_t released;          // staged form: the last epoch whose slots the home rank has emptied
};
struct SlotOffsets { unsigned long long v[16]; };
static_assert(sizeof(Exchange) <= 1024, "ALAC_B200_EXCHANGE_BYTES");
constexpr unsigned long long kExchangeTimeoutNs = 20ull * 1000 * 1000 * 1000;

__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// spin until *flag == epoch; false on timeout
__device__ __forceinline__ bool spin_until(const uint32_t *flag, uint32_t epoch)
{
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys_u32(flag) != epoch) {
        if (global_timer_ns() - t0 > kExchangeTimeoutNs) return false;
        __nanosleep(200);
    }
    return true;
}

// one warp: publish this rank's total (*total_word = offsets[P] of the rank's scan), then sum the totals in front
static __global__ void xchg_publish_resolve_kernel(Exchange *x, uint32_t rank, uint32_t epoch, const uint64_t *total_word, uint64_t *base_out,
                                            uint32_t *err)
{
    const uint32_t slot = epoch & 1u, lane = threadIdx.x;
    if (lane == 0) {
        *reinterpret_cast<volatile unsigned long long *>(&x->total[slot][rank]) = *total_word;
        __threadfence_system();
        st_release_sys_u32(&x->ready[slot][rank], epoch);
    }
    unsigned long long mine = 0;
    bool ok = true;
    if (lane < rank) {
        ok = spin_until(&x->ready[slot][lane], epoch);
        if (ok) mine = ld_relaxed_sys_u64(&x->total[slot][lane]);
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, d);
    const bool all_ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) {
        *base_out = all_ok ? mine : 0ull;
        if (!all_ok) *err = 1u;
    }
}

// after the rank's last assemble launch: its block (and sizes) are in place
static __global__ void xchg_done_kernel(Exchange *x, uint32_t rank, uint32_t epoch)
{
    __threadfence_system();
    st_release_sys_u32(&x->done[epoch & 1u][rank], epoch);
}

// home rank: every rank's block is in place; also totals the job (sum of all ranks' bytes)
static __global__ void xchg_wait_all_kernel(Exchange *x, uint32_t n_ranks, uint32_t epoch, uint64_t *job_total, uint32_t *err)
{
    const uint32_t slot = epoch & 1u, lane = threadIdx.x;
    bool ok = true;
    unsigned long long mine = 0;
    if (lane < n_ranks) {
        ok = spin_until(&x->done[slot][lane], epoch);
        if (ok) mine = ld_relaxed_sys_u64(&x->total[slot][lane]);
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, d);
    const bool all_ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) {
        if (job_total) *job_total = mine;
        if (!all_ok) *err = 2u;
    }
}

// staged form, home rank: every rank's slot is complete -> close the gaps.  Block r (r >= 1) moves from its slot of the
// staging area to the sum of the totals in front of it inside the job's buffer; the destination is written in aligned
// 32-bit words (two aligned source words and a funnel shift each), the ragged ends byte by byte.  Then the slots are
// released for the next epoch.  Runs entirely on the device: the home rank's host thread is free to start decoding.
static __global__ void __launch_bounds__(256) xchg_compact_kernel(Exchange *x, uint32_t n_ranks, uint32_t epoch, SlotOffsets so,
                                                                  const uint8_t *staging, uint8_t *dst, unsigned long long cap, uint32_t *err)
{
    const uint32_t slot = epoch & 1u;
    unsigned long long at = x->total[slot][0];
    const unsigned long long gtid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x, gsize = (unsigned long long)gridDim.x * blockDim.x;
    for (uint32_t r = 1; r < n_ranks; r++) {
        const unsigned long long len = x->total[slot][r];
        if (at + len > cap) { if (gtid == 0) *err = 3u; return; }
        const uint8_t *src = staging + so.v[r];
        uint8_t *d = dst + at;
        // destination in aligned 16-byte vectors; the ragged ends (< 16 bytes each) byte by byte
        const unsigned long long head = min((unsigned long long)((16u - (uint32_t)(reinterpret_cast<uintptr_t>(d) & 15u)) & 15u), len);
        const unsigned long long nvec = (len - head) >> 4;
        const unsigned long long tail0 = head + (nvec << 4);
        if (gtid < head) d[gtid] = src[gtid];
        if (gtid < len - tail0) d[tail0 + gtid] = src[tail0 + gtid];
        // vector i: destination d + head + 16 i; source src + head + 16 i = aligned vector base + k words + sh bits
        const uintptr_t s0 = reinterpret_cast<uintptr_t>(src + head);
        const uint4 *sv = reinterpret_cast<const uint4 *>(s0 & ~(uintptr_t)15);
        const uint32_t k = (uint32_t)(s0 & 15u) >> 2, sh = (uint32_t)(s0 & 3u) * 8u;
        uint4 *dv = reinterpret_cast<uint4 *>(d + head);
        const bool exact = (s0 & 15u) == 0;
        // (when the source is not vector-aligned the last vector reads one source vector past its own: slots are
        //  256-byte aligned and at least 16 bytes of slack follow every block inside the staging area)
        for (unsigned long long i0 = gtid; i0 < nvec; i0 += 4 * gsize) {
            uint4 lo[4], hi[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const unsigned long long i = i0 + (unsigned long long)u * gsize;
                if (i < nvec) { lo[u] = __ldg(sv + i); if (!exact) hi[u] = __ldg(sv + i + 1); }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const unsigned long long i = i0 + (unsigned long long)u * gsize;
                if (i >= nvec) continue;
                uint4 o = lo[u];
                if (!exact) {
                    const uint32_t w[9] = {lo[u].x, lo[u].y, lo[u].z, lo[u].w, hi[u].x, hi[u].y, hi[u].z, hi[u].w, 0u};
                    uint32_t q[5];
#pragma unroll
                    for (int j = 0; j < 5; j++) q[j] = k == 0 ? w[j] : k == 1 ? w[j + 1] : k == 2 ? w[j + 2] : w[j + 3];
                    o.x = __funnelshift_r(q[0], q[1], sh); o.y = __funnelshift_r(q[1], q[2], sh);
                    o.z = __funnelshift_r(q[2], q[3], sh); o.w = __funnelshift_r(q[3], q[4], sh);
                }
                dv[i] = o;
            }
        }
        at += len;
    }
}
static __global__ void xchg_release_kernel(Exchange *x, uint32_t epoch)
{
    __threadfence_system();
    st_release_sys_u32(&x->released, epoch);
}
// a rank may only refill its slot once the home rank has emptied it (the previous epoch's compaction)
static __global__ void xchg_wait_released_kernel(Exchange *x, uint32_t want, uint32_t *err)
{
    const unsigned long long t0 = global_timer_ns();
    while ((int32_t)(ld_acquire_sys_u32(&x->released) - want) < 0) {
        if (global_timer_ns() - t0 > kExchangeTimeoutNs) { *err = 4u; return; }
        __nanosleep(200);
    }
}

// ---- packet assembly --------------------------------------------------------------------------------------
enum : uint32_t { REG_WORDS = 0, REG_SHIFT = 1, REG_RAW = 2 };
struct Region {
    uint32_t dst;       // first bit inside the packet
    uint32_t nbits;
    uint32_t kind;
    uint32_t elem;      // element slot (REG_SHIFT / REG_RAW)
    const uint32_t *words;   // REG_WORDS: MSB-first words (shared or global)
};

__device__ __forceinline__ uint32_t bits_from_words(const uint32_t *w, uint32_t off, uint32_t nb)
{
    // nb 1..32 bits starting at bit `off` of an MSB-first word array (reads w[i] and, if needed, w[i+1])
    const uint32_t i = off >> 5, sh = off & 31u;
    const uint32_t hi = w[i];
    const uint32_t lo = (sh + nb > 32u) ? w[i + 1] : 0u;
    return __funnelshift_l(lo, hi, sh) >> (32u - nb);
}

// fixed-width entry arrays (shift region, escape samples): bits [off, off+nb) of the
// concatenation of W-bit entries
template <class EntryFn>
__device__ __forceinline__ uint32_t bits_from_entries(uint32_t W, uint32_t off, uint32_t nb, EntryFn entry)
{
    uint32_t i = off / W;
    const uint32_t bo = off - i * W;
    uint64_t acc = 0;
    uint32_t filled = 0;
    while (filled < bo + nb) {
        acc = (W == 32 ? (acc << 32) : (acc << W)) | (uint64_t)entry(i++);
        filled += W;
    }
    const uint64_t v = acc >> (filled - bo - nb);
    return nb == 32 ? (uint32_t)v : ((uint32_t)v & ((1u << nb) - 1u));
}

struct AsmArgs {
    const uint8_t *pcm;
    const uint64_t *pkt_frame;
    const uint32_t *pkt_samples;
    const ElemRec *recs;
    const uint32_t *scratch;
    uint32_t cap_words;
    const uint32_t *sizes;
    const uint64_t *offsets;
    uint8_t *out;                 // local memory, or another GPU's (peer stores over NVLink)
    const uint64_t *base;         // optional device word: byte offset of this launch's block inside `out`
    uint32_t pkt_base;            // chunk: first packet; recs / scratch are chunk-relative
    uint32_t num_packets;         // chunk: packets handled by this launch
    EncLayout lay;
};

constexpr int kAsmWarps = 4;
constexpr uint32_t kAsmStageBytes = 3136;   // PCM span of 128 output words of a shift region: 24-bit stereo 1.5 KB, 24-bit mono in a stereo-wide frame 3.1 KB
constexpr int kHdrWords = 13;       // 7 + 16 + 32 + 16 + 2*(16+128) = 359 bits
constexpr int kMaxRegions = 8 * 4 + 1;

template <int DEPTH>
__global__ void __launch_bounds__(kAsmWarps * 32) enc_assemble_kernel(AsmArgs A)
{
    __shared__ uint32_t s_hdr[kAsmWarps][8][kHdrWords];
    __shared__ Region s_reg[kAsmWarps][kMaxRegions];
    __shared__ uint32_t s_nreg[kAsmWarps];
    __shared__ uint32_t s_end_word[kAsmWarps];
    __shared__ __align__(16) uint8_t s_stage[DepthTraits<DEPTH>::kShift ? kAsmWarps : 1][DepthTraits<DEPTH>::kShift ? kAsmStageBytes : 16];

    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint32_t rel = blockIdx.x * kAsmWarps + wid;      // chunk-relative packet
    if (rel >= A.num_packets) return;
    const uint32_t pkt = A.pkt_base + rel;
    constexpr uint32_t bps = DepthTraits<DEPTH>::kBytes;
    constexpr uint32_t shift = DepthTraits<DEPTH>::kShift;
    const uint32_t n = A.pkt_samples[pkt];
    const uint32_t partial = (n != A.lay.frame_size);
    const uint32_t E = A.lay.elems_per_packet;
    const uint32_t stride = A.lay.channels * bps;
    const uint8_t *frame_base = A.pcm + A.pkt_frame[pkt] * A.lay.channels * bps;

    // lanes 0..E-1 build their element's header words
    if (lane < E) {
        const ElemRec &r = A.recs[(size_t)rel * E + lane];
        uint32_t *h = s_hdr[wid][lane];
        BitSink bs;
        bs.start(h, kHdrWords);
        const uint32_t tag = A.lay.elem_tag[lane];
        const bool stereo = (tag == ID_CPE);
        bs.put(tag, 3);
        bs.put(A.lay.elem_inst[lane], 4);
        bs.put(0, 12);
        if (r.escape) {
            bs.put((partial << 3) | 1u, 4);                                       // :761-765
            if (partial) bs.put(n, 32);
        } else {
            bs.put((partial << 3) | ((shift / 8) << 1), 4);                       // :466-471
            if (partial) bs.put(n, 32);
            bs.put(stereo ? (uint32_t)kMixBits : 0u, 8);
            bs.put(stereo ? (uint32_t)r.mix_res : 0u, 8);
            bs.put((0u << 4) | kDenShift, 8);                                     // :477-485
            bs.put((4u << 5) | r.num_u, 8);
            for (uint32_t i = 0; i < r.num_u; i++) bs.put((uint16_t)r.coef_u[i], 16);
            if (stereo) {
                bs.put((0u << 4) | kDenShift, 8);
                bs.put((4u << 5) | r.num_v, 8);
                for (uint32_t i = 0; i < r.num_v; i++) bs.put((uint16_t)r.coef_v[i], 16);
            }
        }
        bs.finish();
    }
    __syncwarp();
    // lane 0 lays out the regions
    if (lane == 0) {
        uint32_t nr = 0, at = 0;
        Region *R = s_reg[wid];
        for (uint32_t e = 0; e < E; e++) {
            const ElemRec &r = A.recs[(size_t)rel * E + e];
            const bool stereo = (A.lay.elem_tag[e] == ID_CPE);
            const uint32_t nch = stereo ? 2u : 1u;
            uint32_t hb;
            if (r.escape) hb = 7 + 16 + (partial ? 32u : 0u);
            else hb = 7 + 16 + (partial ? 32u : 0u) + 16 + (16 + 16 * r.num_u) + (stereo ? (16 + 16 * r.num_v) : 0u);
            R[nr++] = Region{at, hb, REG_WORDS, e, s_hdr[wid][e]};
            at += hb;
            if (r.escape) {
                if (n) { R[nr++] = Region{at, n * nch * DEPTH, REG_RAW, e, nullptr}; at += n * nch * DEPTH; }
            } else {
                if (shift && n) { R[nr++] = Region{at, n * nch * shift, REG_SHIFT, e, nullptr}; at += n * nch * shift; }
                const uint32_t *su = A.scratch + ((size_t)rel * A.lay.chains_per_packet + A.lay.elem_chain[e]) * A.cap_words;
                if (r.bits_u) { R[nr++] = Region{at, r.bits_u, REG_WORDS, e, su}; at += r.bits_u; }
                if (stereo && r.bits_v) { R[nr++] = Region{at, r.bits_v, REG_WORDS, e, su + A.cap_words}; at += r.bits_v; }
            }
        }
        // ID_END (3 bits of ones); padding bits stay zero
        s_end_word[wid] = 0xE0000000u;
        R[nr++] = Region{at, 3, REG_WORDS, 0, &s_end_word[wid]};
        s_nreg[wid] = nr;
    }
    __syncwarp();

    const uint32_t nreg = s_nreg[wid];
    const Region *R = s_reg[wid];
    const uint32_t size = A.sizes[pkt];
    uint8_t *dst = A.out + (A.base ? *A.base : 0ull) + A.offsets[pkt];
    // output is written as 4-byte words aligned in the OUTPUT buffer; a = bytes before the first aligned word
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3u);
    const uint32_t lead_bytes = mis ? (4u - mis) : 0u;     // bytes of the packet before the first aligned word
    // word g (g >= 0) covers packet bytes [lead + 4g, lead + 4g + 4); g = -1 is the leading partial word
    const uint32_t words_total = size > lead_bytes ? (size - lead_bytes + 3) / 4 : 0u;
    uint32_t cursor = 0;
    // gather the bits [bit0, bit0 + 8*nbytes) of the packet into an MSB-first word
    auto gather = [&](uint32_t byte0, uint32_t nbytes) -> uint32_t {
        const uint32_t bit0 = byte0 * 8u, bit1 = bit0 + nbytes * 8u;
        uint32_t word = 0;
        while (cursor < nreg && R[cursor].dst + R[cursor].nbits <= bit0) cursor++;
        for (uint32_t ri = cursor; ri < nreg && R[ri].dst < bit1; ri++) {
            const Region &r = R[ri];
            const uint32_t lo = max(bit0, r.dst), hi = min(bit1, r.dst + r.nbits);
            const uint32_t off = lo - r.dst, nb = hi - lo;
            uint32_t v;
            if (r.kind == REG_WORDS) {
                v = bits_from_words(r.words, off, nb);
            } else {
                const bool stereo = (A.lay.elem_tag[r.elem] == ID_CPE);
                const uint8_t *eb = frame_base + (size_t)A.lay.elem_chan[r.elem] * bps;
                if (r.kind == REG_SHIFT) {
                    // :488-500 stereo (loL << s) | loR in 2s bits; :934-938 mono lo in s bits
                    constexpr uint32_t sh = shift ? shift : 8u;
                    constexpr uint32_t mask = (1u << sh) - 1u;
                    if (stereo) {
                        v = bits_from_entries(2 * sh, off, nb, [&](uint32_t i) -> uint32_t {
                            if (i >= n) return 0u;
                            const uint8_t *p = eb + (size_t)i * stride;
                            const uint32_t l = load_raw_bits<DEPTH>(p) & mask, rr = load_raw_bits<DEPTH>(p + bps) & mask;
                            return (l << sh) | rr;
                        });
                    } else {
                        v = bits_from_entries(sh, off, nb, [&](uint32_t i) -> uint32_t {
                            if (i >= n) return 0u;
                            return load_raw_bits<DEPTH>(eb + (size_t)i * stride) & mask;
                        });
                    }
                } else {
                    const uint32_t nch = stereo ? 2u : 1u;
                    v = bits_from_entries(DEPTH, off, nb, [&](uint32_t i) -> uint32_t {
                        if (i >= n * nch) return 0u;
                        const uint32_t s = stereo ? (i >> 1) : i, c = stereo ? (i & 1u) : 0u;
                        return load_raw_bits<DEPTH>(eb + (size_t)s * stride + c * bps);
                    });
                }
            }
            word |= v << (bit0 + 32u - hi);     // word holds packet bits [bit0, bit0+32) MSB-first
        }
        return word;
    };
    // bytes before the first 4-byte-aligned output address: lane 0, byte stores
    if (lane == 0 && lead_bytes) {
        const uint32_t nb = min(lead_bytes, size);
        const uint32_t word = gather(0, nb);
        for (uint32_t b = 0; b < nb; b++) dst[b] = (uint8_t)(word >> (24 - 8 * b));
    }
    auto put_word = [&](uint32_t g) {
        const uint32_t byte0 = lead_bytes + 4u * g;
        const uint32_t nbytes = min(4u, size - byte0);
        const uint32_t word = gather(byte0, nbytes);
        if (nbytes == 4) {
            *reinterpret_cast<uint32_t *>(dst + byte0) = bswap32(word);
        } else {
            for (uint32_t b = 0; b < nbytes; b++) dst[byte0 + b] = (uint8_t)(word >> (24 - 8 * b));
        }
    };
    // Almost every output word lies wholly inside ONE long region -- a Golomb stream in the slab or the shift bytes of
    // the PCM -- and is then a funnel shift of two consecutive source words (or two or three fixed-width PCM entries):
    // those words take the short path below, region by region.  The few words in between (headers, region
    // boundaries, the tail of the packet) go through the general gather.
    const uint32_t lead_bits = lead_bytes * 8u;
    const uint32_t full_words = size > lead_bytes ? (size - lead_bytes) / 4u : 0u;      // output words that are whole
    uint32_t done_to = 0;                                                               // words [0, done_to) are written
    for (uint32_t ri = 0; ri < nreg; ri++) {
        const Region r = R[ri];
        if (r.nbits < 96u || r.kind == REG_RAW || (r.kind == REG_WORDS && r.dst + r.nbits <= lead_bits)) continue;
        // whole output words inside the region: word g = packet bits [lead_bits + 32 g, + 32)
        const uint32_t g_lo = r.dst > lead_bits ? (r.dst - lead_bits + 31u) / 32u : 0u;
        const uint32_t g_hi = min(full_words, (r.dst + r.nbits - lead_bits) / 32u);
        if (g_hi <= g_lo || g_lo < done_to) continue;
        for (uint32_t g = done_to + lane; g < g_lo; g += 32) put_word(g);                    // the words before it
        // four words per lane and trip: all their loads are issued before the first store (the kernel is otherwise
        // bound by the latency of loads that are used at once)
        constexpr uint32_t kU = 4;
        if (r.kind == REG_WORDS) {
            for (uint32_t g0 = g_lo + lane; g0 < g_hi; g0 += 32u * kU) {
                uint32_t whi[kU], wlo[kU];
#pragma unroll
                for (uint32_t u = 0; u < kU; u++) {
                    const uint32_t g = min(g0 + 32u * u, g_hi - 1u);
                    const uint32_t i = (lead_bits + 32u * g - r.dst) >> 5;
                    whi[u] = r.words[i];
                    wlo[u] = r.words[i + 1];        // (the slab has spare words past every stream)
                }
#pragma unroll
                for (uint32_t u = 0; u < kU; u++) {
                    const uint32_t g = g0 + 32u * u;
                    if (g < g_hi) *reinterpret_cast<uint32_t *>(dst + lead_bytes + 4u * g) = bswap32(__funnelshift_l(wlo[u], whi[u], (lead_bits - r.dst) & 31u));
                }
            }
        } else {
            constexpr uint32_t sh = shift ? shift : 8u;
            constexpr uint32_t mask = (1u << sh) - 1u;
            const bool stereo = (A.lay.elem_tag[r.elem] == ID_CPE);
            const uint8_t *eb = frame_base + (size_t)A.lay.elem_chan[r.elem] * bps;
            // The shift bytes sit at a stride of one sample-frame in the PCM (6 bytes for 24-bit stereo): fetched straight
            // from global memory that is two to six byte loads per output word.  When the element is (nearly) the whole
            // frame, the warp instead copies the PCM span of 128 output words into shared memory with aligned 16-byte
            // loads (an aligned vector that holds one valid byte lies inside a mapped granule) and picks the bytes there.
            const uint32_t Wb = stereo ? 2u * sh : sh;
            const uint32_t elem_bytes = (stereo ? 2u : 1u) * bps;
            uint32_t g_from = g_lo;
            if (shift != 0 && ((128u * 32u) / Wb + 2u) * stride + elem_bytes + 32u <= kAsmStageBytes) {
                uint8_t *stage = s_stage[wid];
                for (uint32_t t0 = g_lo; t0 < g_hi; t0 += 32u * kU) {
                    const uint32_t gw = min(32u * kU, g_hi - t0);
                    const uint32_t e_lo = (lead_bits + 32u * t0 - r.dst) / Wb;
                    const uint32_t e_hi = (lead_bits + 32u * (t0 + gw) - r.dst - 1u) / Wb;
                    const uint8_t *p0 = eb + (size_t)e_lo * stride;
                    const uint32_t lead = (uint32_t)(reinterpret_cast<uintptr_t>(p0) & 15u);
                    const uint4 *a0 = reinterpret_cast<const uint4 *>(p0 - lead);
                    const uint32_t nvec = (lead + (e_hi - e_lo) * stride + elem_bytes + 15u) >> 4;
                    __syncwarp();
                    for (uint32_t v = lane; v < nvec; v += 32) reinterpret_cast<uint4 *>(stage)[v] = __ldg(a0 + v);
                    __syncwarp();
                    const uint8_t *sb = stage + lead;
                    uint32_t word[kU];
#pragma unroll
                    for (uint32_t u = 0; u < kU; u++) {
                        const uint32_t g = min(t0 + lane + 32u * u, g_hi - 1u);
                        const uint32_t off = lead_bits + 32u * g - r.dst;
                        auto low = [&](const uint8_t *q) -> uint32_t {      // the sample's low `sh` bits (little-endian container)
                            return sh == 8 ? (uint32_t)q[0] : ((uint32_t)q[0] | ((uint32_t)q[1] << 8));
                        };
                        if (stereo) {
                            word[u] = bits_from_entries(2 * sh, off, 32u, [&](uint32_t i) -> uint32_t {
                                const uint8_t *q = sb + (i - e_lo) * stride;
                                return (low(q) << sh) | low(q + bps);
                            });
                        } else {
                            word[u] = bits_from_entries(sh, off, 32u, [&](uint32_t i) -> uint32_t { return low(sb + (i - e_lo) * stride); });
                        }
                    }
#pragma unroll
                    for (uint32_t u = 0; u < kU; u++) {
                        const uint32_t g = t0 + lane + 32u * u;
                        if (g < g_hi) *reinterpret_cast<uint32_t *>(dst + lead_bytes + 4u * g) = bswap32(word[u]);
                    }
                }
                g_from = g_hi;
            }
            for (uint32_t g0 = g_from + lane; g0 < g_hi; g0 += 32u * kU) {
                uint32_t word[kU];
#pragma unroll
                for (uint32_t u = 0; u < kU; u++) {
                    const uint32_t g = min(g0 + 32u * u, g_hi - 1u);
                    const uint32_t off = lead_bits + 32u * g - r.dst;
                    if (stereo) {
                        word[u] = bits_from_entries(2 * sh, off, 32u, [&](uint32_t i) -> uint32_t {
                            const uint8_t *p = eb + (size_t)i * stride;
                            return ((load_raw_bits<DEPTH>(p) & mask) << sh) | (load_raw_bits<DEPTH>(p + bps) & mask);
                        });
                    } else {
                        word[u] = bits_from_entries(sh, off, 32u, [&](uint32_t i) -> uint32_t { return load_raw_bits<DEPTH>(eb + (size_t)i * stride) & mask; });
                    }
                }
#pragma unroll
                for (uint32_t u = 0; u < kU; u++) {
                    const uint32_t g = g0 + 32u * u;
                    if (g < g_hi) *reinterpret_cast<uint32_t *>(dst + lead_bytes + 4u * g) = bswap32(word[u]);
                }
            }
        }
        done_to = g_hi;
    }
    for (uint32_t g = done_to + lane; g < words_total; g += 32) put_word(g);
}


// ---- host-side launchers ------------------------------------------------------------------------------------------
// The kernels above are templates on the bit depth; each depth is instantiated in its own translation unit
// (alac_kernels_d16.cu ...), so the four compile in parallel and the engine's translation unit holds no kernel code.
// ev: optional events recorded around the kernels -- {before, between search and final, after} for the pair launch,
// then the same three for the mono launch (split form only).  Returns the number of kernels launched.
template <int DEPTH, bool PACKED, bool WRAP>
static uint32_t enc_launch_search_v(cudaStream_t s, const EncArgs &A, uint32_t mono_mask, uint32_t pair_mask, const JobLists *split, cudaEvent_t *ev,
                                    int dense)
{
    const uint32_t pairs = __builtin_popcount(pair_mask), monos = __builtin_popcount(mono_mask);
    uint32_t launches = 0;
    if (pairs) {
        const uint64_t threads = (uint64_t)A.num_segments * pairs * 2;
        if (split) {
            const uint32_t ctas = (uint32_t)((threads + 31) / 32);
            if (ev) cudaEventRecord(ev[0], s);
            if (PACKED && DEPTH != 16 && dense) enc_search_split_kernel<DEPTH, true, PACKED && DEPTH != 16, WRAP, PACKED && DEPTH != 16><<<ctas, 32, 0, s>>>(A, pairs, pair_mask, *split);
            else enc_search_split_kernel<DEPTH, true, PACKED, WRAP><<<ctas, 32, 0, s>>>(A, pairs, pair_mask, *split);
            if (ev) cudaEventRecord(ev[1], s);
            if (dense == 2) enc_final2_kernel<DEPTH, true, WRAP, true><<<2 * ctas, 64, 0, s>>>(A, *split, ctas);
            else if (dense) enc_final2_kernel<DEPTH, true, WRAP, false><<<2 * ctas, 32, 0, s>>>(A, *split, ctas);
            else enc_final_kernel<DEPTH, true, PACKED, WRAP><<<2 * ctas, 32, 0, s>>>(A, *split, ctas);
            if (ev) cudaEventRecord(ev[2], s);
            launches += 2;
        } else {
            enc_search_kernel<DEPTH, true, PACKED, WRAP>
                <<<(uint32_t)((threads + kChainThreads - 1) / kChainThreads), kSearchThreads, 0, s>>>(A, pairs, pair_mask);
            launches += 1;
        }
    }
    if (monos) {
        const uint64_t threads = (uint64_t)A.num_segments * monos;
        if (split) {
            const uint32_t ctas = (uint32_t)((threads + 31) / 32);
            JobLists Qm = *split;
            Qm.counts += 2;     // the mono launch has its own pair of counters
            if (ev) cudaEventRecord(ev[3], s);
            if (dense) enc_search_split_kernel<DEPTH, false, false, WRAP, true><<<ctas, 32, 0, s>>>(A, monos, mono_mask, Qm);
            else enc_search_split_kernel<DEPTH, false, false, WRAP><<<ctas, 32, 0, s>>>(A, monos, mono_mask, Qm);
            if (ev) cudaEventRecord(ev[4], s);
            if (dense == 2) enc_final2_kernel<DEPTH, false, WRAP, true><<<2 * ctas, 64, 0, s>>>(A, Qm, ctas);
            else if (dense) enc_final2_kernel<DEPTH, false, WRAP, false><<<2 * ctas, 32, 0, s>>>(A, Qm, ctas);
            else enc_final_kernel<DEPTH, false, false, WRAP><<<2 * ctas, 32, 0, s>>>(A, Qm, ctas);
            if (ev) cudaEventRecord(ev[5], s);
            launches += 2;
        } else {
            enc_search_kernel<DEPTH, false, false, WRAP>
                <<<(uint32_t)((threads + kChainThreads - 1) / kChainThreads), kSearchThreads, 0, s>>>(A, monos, mono_mask);
            launches += 1;
        }
    }
    return launches;
}

// packed: pure stereo PCM at 8-byte alignment (one wide load per sample-frame);
// wrap: the int16 coefficient range could be left during a segment, so every update re-wraps;
// dense: 1 or 2 = a mono or stereo stream (the element is the whole sample-frame) with every packet on a 4-byte boundary: the
//        split form then runs its final pass through the word ring, 2 = as the two-warp kernel (launches that do not fill the GPU)
template <int DEPTH>
uint32_t enc_launch_search(cudaStream_t s, const EncArgs &A, uint32_t mono_mask, uint32_t pair_mask, bool packed, bool wrap,
                           const JobLists *split, cudaEvent_t *ev, int dense)
{
    if (packed) return wrap ? enc_launch_search_v<DEPTH, true, true>(s, A, mono_mask, pair_mask, split, ev, dense)
                            : enc_launch_search_v<DEPTH, true, false>(s, A, mono_mask, pair_mask, split, ev, dense);
    return wrap ? enc_launch_search_v<DEPTH, false, true>(s, A, mono_mask, pair_mask, split, ev, dense)
                : enc_launch_search_v<DEPTH, false, false>(s, A, mono_mask, pair_mask, split, ev, dense);
}

template <int DEPTH>
void enc_launch_assemble(cudaStream_t s, const AsmArgs &A)
{
    enc_assemble_kernel<DEPTH><<<(A.num_packets + kAsmWarps - 1) / kAsmWarps, kAsmWarps * 32, 0, s>>>(A);
}

#ifndef ALAC_INSTANTIATE_DEPTH
#define ALAC_ENC_EXTERN(D)                                                                                                            \
    extern template uint32_t enc_launch_search<D>(cudaStream_t, const EncArgs &, uint32_t, uint32_t, bool, bool, const JobLists *, cudaEvent_t *, int); \
    extern template void enc_launch_assemble<D>(cudaStream_t, const AsmArgs &);
ALAC_ENC_EXTERN(16) ALAC_ENC_EXTERN(20) ALAC_ENC_EXTERN(24) ALAC_ENC_EXTERN(32)
#undef ALAC_ENC_EXTERN
#endif

}  // namespace alacb
