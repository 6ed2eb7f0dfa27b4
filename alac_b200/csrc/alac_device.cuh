// alac_device.cuh -- device-side ALAC primitives shared by the encode and decode kernels.
//
// Everything here is exact integer arithmetic; the semantics follow the reference's host
// primitives (file:line relative to /root/reference):
//   predictor      codec/dp_enc.c:77-388 (pc_block), codec/dp_dec.c:55-381 (unpc_block)
//   Golomb coder   codec/ag_enc.c:249-367 (dyn_comp), codec/ag_dec.c:272-362 (dyn_decomp)
// but are re-shaped for one-chain-per-lane streaming: each primitive consumes / produces ONE
// sample per call so predictor, entropy coder and (un)mixing fuse into a single pass with no
// intermediate arrays.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace alacb {

// ---- constants (codec/aglib.h:36-55, codec/dplib.h:40-45) ----------------------------------
constexpr uint32_t kQbShift = 9;
constexpr uint32_t kQb = 1u << kQbShift;
constexpr uint32_t kPb0 = 40, kMb0 = 10, kKb0 = 14;
constexpr uint32_t kMaxPrefix = 9;
constexpr uint32_t kRunRawBits = 16;
constexpr uint32_t kMeanClamp = 0xffffu;
constexpr uint32_t kDenShift = 9;
constexpr int kMaxRes = 4;
constexpr int kMixBits = 2;

enum : uint32_t { ID_SCE = 0, ID_CPE = 1, ID_CCE = 2, ID_LFE = 3, ID_DSE = 4, ID_PCE = 5, ID_FIL = 6, ID_END = 7 };

// ---- tiny helpers ---------------------------------------------------------------------------
__device__ __forceinline__ int32_t sext_bits(int32_t v, uint32_t chanshift)
{
    // "(x << chanshift) >> chanshift", codec/dp_enc.c:232
    return (int32_t)((uint32_t)v << chanshift) >> chanshift;
}
__device__ __forceinline__ int32_t sext16(int32_t v) { return (int32_t)(int16_t)v; }
__device__ __forceinline__ int32_t sign3(int32_t v) { return min(max(v, -1), 1); }
__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }
__device__ __forceinline__ uint32_t bfind_u32(uint32_t v)     // index of the highest set bit = 31 - clz(v), v != 0
{
    uint32_t r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}


// ---- PCM sample access (packed little-endian) ------------------------------------------------
// full-width, right-aligned, sign-extended sample; codec/matrix_enc.cu:72-99,120-159,186-282,330-391
template <int DEPTH>
__device__ __forceinline__ int32_t load_sample(const uint8_t *p)
{
    if (DEPTH == 16) {
        return (int32_t)__ldg(reinterpret_cast<const int16_t *>(p));
    } else if (DEPTH == 32) {
        return __ldg(reinterpret_cast<const int32_t *>(p));
    } else {
        uint32_t w = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16);
        return DEPTH == 20 ? ((int32_t)(w << 8) >> 12) : ((int32_t)(w << 8) >> 8);
    }
}
// raw container bits of one sample, as the escape path writes them (codec/ALACEncoder.cu:761-803)
template <int DEPTH>
__device__ __forceinline__ uint32_t load_raw_bits(const uint8_t *p)
{
    if (DEPTH == 16) return (uint32_t)__ldg(reinterpret_cast<const uint16_t *>(p));
    if (DEPTH == 32) return (uint32_t)__ldg(reinterpret_cast<const int32_t *>(p));
    uint32_t w = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16);
    return DEPTH == 20 ? (w >> 4) : w;
}
template <int DEPTH> struct DepthTraits {
    static constexpr uint32_t kBytes = DEPTH == 16 ? 2 : DEPTH == 32 ? 4 : 3;
    static constexpr uint32_t kShift = DEPTH == 32 ? 16 : DEPTH == 24 ? 8 : 0;   // codec/ALACEncoder.cu:327-332
};
template <int DEPTH>
__device__ __forceinline__ void store_sample(uint8_t *p, int32_t v)
{
    // codec/ALACDecoder.cu:193-495 output packing
    if (DEPTH == 16) {
        *reinterpret_cast<int16_t *>(p) = (int16_t)v;
    } else if (DEPTH == 32) {
        *reinterpret_cast<int32_t *>(p) = v;
    } else {
        if (DEPTH == 20) v = (int32_t)((uint32_t)v << 4);
        p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16);
    }
}

// ---- sign-LMS predictor step -------------------------------------------------------------------
// hist[0] = newest previous sample ... hist[TAPS-1], hist[TAPS] = "top" (codec/dp_enc.c:202-214).
// Coefficients are carried as int32 holding int16 values.  WRAP = true re-wraps every update to
// int16 exactly like the reference's int16_t registers; WRAP = false is used when the caller has
// proved the values cannot leave the int16 range during the pass (|a| + steps <= 32767), in which
// case the two are identical.
template <int TAPS, bool WRAP>
__device__ __forceinline__ void lms_adapt(int32_t (&a)[TAPS], const int32_t (&b)[TAPS], int32_t err)
{
    // codec/dp_enc.c:236-329.  Branch-free ladder: the walk goes from the last tap to the first and
    // stops once the running error reaches or crosses zero.
    //   err > 0:  sgn = sign(b);  a -= sgn;  del0 -= w * (( sgn * b) >> 9);  go on while del0 > 0
    //   err < 0:  sgn = sign(b);  a += sgn;  del0 -= w * ((-sgn * b) >> 9);  go on while del0 < 0
    // sgn * b = |b|, and the arithmetic shift of -|b| rounds toward -inf (dp_enc.c:288): (-|b|) >> 9 = -((|b| + 511) >> 9).
    // Both signs therefore are ONE walk on left = |del0|:  left -= w * ((|b| + c) >> 9) with c = 0 or 511, go on while
    // left > 0.  left never grows, so "tap k is reached" is simply left > 0 at that point (no chain of conditions), and
    // the update is a predicated multiply-add.  Per tap: two min/max, two IMAD, a shift, a compare and the predicated
    // IMAD -- the integer ALU pipe and the IMAD pipe issue every other cycle each, so the instruction count and its split
    // over the two pipes is what bounds these kernels.
    const int32_t m = err >> 31;
    const int32_t nsg = -2 * m - 1;                        // +1 for err < 0, -1 otherwise
    const int32_t c = m & ((1 << kDenShift) - 1);
    int32_t left = abs(err);                                // |err| < 2^25: err is a sign-extended chan_bits value
#pragma unroll
    for (int k = TAPS - 1; k >= 0; k--) {
        const int32_t sb = sign3(b[k]);
        asm("{\n\t.reg .pred p;\n\tsetp.gt.s32 p, %3, 0;\n\t@p mad.lo.s32 %0, %1, %2, %0;\n\t}" : "+r"(a[k]) : "r"(sb), "r"(nsg), "r"(left));
        if (WRAP) a[k] = sext16(a[k]);
        left -= (TAPS - k) * ((sb * b[k] + c) >> kDenShift);
    }
}

// encode: returns the residual of x given the history, adapts coefficients, shifts the history in
template <int TAPS, bool WRAP>
__device__ __forceinline__ int32_t predict_enc_step(int32_t x, int32_t (&hist)[TAPS + 1], int32_t (&a)[TAPS], uint32_t chanshift)
{
    const int32_t top = hist[TAPS];
    int32_t b[TAPS];
    int32_t acc = 1 << (kDenShift - 1);
#pragma unroll
    for (int k = 0; k < TAPS; k++) {
        b[k] = top - hist[k];
        acc -= a[k] * b[k];
    }
    const int32_t err = sext_bits(x - top - (acc >> kDenShift), chanshift);     // dp_enc.c:228-233
    lms_adapt<TAPS, WRAP>(a, b, err);
#pragma unroll
    for (int k = TAPS; k > 0; k--) hist[k] = hist[k - 1];
    hist[0] = x;
    return err;
}

// decode: returns the reconstructed sample for residual err (codec/dp_dec.c:206-282)
template <int TAPS, bool WRAP>
__device__ __forceinline__ int32_t predict_dec_step(int32_t err, int32_t (&hist)[TAPS + 1], int32_t (&a)[TAPS], uint32_t chanshift)
{
    const int32_t top = hist[TAPS];
    int32_t b[TAPS];
    int32_t acc = 1 << (kDenShift - 1);
#pragma unroll
    for (int k = 0; k < TAPS; k++) {
        b[k] = top - hist[k];
        acc -= a[k] * b[k];
    }
    const int32_t x = sext_bits(err + top + (acc >> kDenShift), chanshift);
    lms_adapt<TAPS, WRAP>(a, b, err);
#pragma unroll
    for (int k = TAPS; k > 0; k--) hist[k] = hist[k - 1];
    hist[0] = x;
    return x;
}

// ---- adaptive Golomb: streaming encoder state ------------------------------------------------------
// dyn_comp consumes a block; here the same state machine is advanced one residual at a time.
struct AgEnc {
    uint32_t mb;        // running mean, codec/ag_enc.c:273
    uint32_t zmode;     // 1 right after a zero run
    uint32_t in_run;    // counting zeros (codec/ag_enc.c:328-350)
    uint32_t nz;        // zeros counted so far
    uint32_t c;         // samples consumed
    uint32_t count;     // block length
    uint32_t bits;      // bits produced
    __device__ __forceinline__ void start(uint32_t n)
    {
        mb = kMb0; zmode = 0; in_run = 0; nz = 0; c = 0; count = n; bits = 0;
    }
};

// MSB-first bit sink: 32-bit words, word w holds stream bits [32w, 32w+32) with bit 32w in the MSB
// The destination must hold the worst case: every sample code is <= 9 + 23 = 32 bits and a run code
// (<= 25 bits) only ever follows a short sample code, so frame_size + 1 words always suffice.
struct BitSink {
    uint32_t *dst;      // next word to write
    uint64_t acc;
    uint32_t nacc;
    __device__ __forceinline__ void start(uint32_t *p, uint32_t /*cap_words*/) { dst = p; acc = 0; nacc = 0; }
    __device__ __forceinline__ void put(uint32_t value, uint32_t len)   // len 1..32, value < 2^len
    {
        acc = (acc << len) | value;
        nacc += len;
        if (nacc >= 32) {
            nacc -= 32;
            *dst++ = (uint32_t)(acc >> nacc);
        }
    }
    __device__ __forceinline__ void finish()
    {
        if (nacc) *dst++ = (uint32_t)(acc << (32 - nacc));
    }
};
struct NoSink {
    __device__ __forceinline__ void put(uint32_t, uint32_t) {}
};

// exact n / (2^k - 1) for n < 9 * (2^k - 1): floor(2^32 / m) + 1, k = 0..15 (k = 0 unused)
static __constant__ uint32_t c_div_magic[16] = {
    0u, 0u /* m = 1 handled apart */, 1431655766u, 613566757u, 286331154u, 138547333u, 68174085u, 33818641u,
    16843010u, 8405025u, 4198405u, 2098178u, 1048833u, 524353u, 262161u, 131077u
};

// codec/ag_enc.c:115-148 dyn_code for a zero-run length: (length, code word) from the running mean and the run.
// Out of line on purpose: runs are rare, the routine holds an integer division, and ag_put is inlined into every
// unrolled predictor loop -- keeping this body out of those loops keeps them inside the instruction cache.
__device__ __forceinline__ uint2 ag_run_code_inline(uint32_t mb, uint32_t n)
{
    const uint32_t k = (uint32_t)__clz((int)mb) - 24u + ((mb + 16u) >> 6);   // ag_enc.c:352
    const uint32_t mz = ((1u << k) - 1u) & ((1u << kKb0) - 1u);
    const uint32_t div = n / mz;
    uint32_t len, value;
    if (div < kMaxPrefix) {
        const uint32_t mod = n - div * mz;
        const uint32_t de = (mod == 0);
        len = div + k + 1 - de;
        value = (((1u << div) - 1u) << (len - div)) + mod + 1 - de;
        if (len > kMaxPrefix + kRunRawBits) { len = kMaxPrefix + kRunRawBits; value = (((1u << kMaxPrefix) - 1u) << kRunRawBits) + n; }
    } else {
        len = kMaxPrefix + kRunRawBits;
        value = (((1u << kMaxPrefix) - 1u) << kRunRawBits) + n;
    }
    return make_uint2(len, value);
}

static __device__ __noinline__ uint2 ag_run_code(uint32_t mb, uint32_t n) { return ag_run_code_inline(mb, n); }

template <bool EMIT, class Sink>
__device__ __forceinline__ void ag_flush_run(AgEnc &s, Sink &sink, uint32_t zmode_after)
{
    // the emitting loops (one per kernel) keep it inline; the many costing loops of the search kernel call it
    const uint2 code = EMIT ? ag_run_code_inline(s.mb, s.nz) : ag_run_code(s.mb, s.nz);
    s.bits += code.x;
    if (EMIT) sink.put(code.y, code.x);
    s.mb = 0;
    s.zmode = zmode_after;
    s.in_run = 0;
}

// one residual through dyn_comp's loop body (codec/ag_enc.c:277-361); encoder-side pb/kb/mb0 are
// always 40/14/10 (codec/ALACEncoder.cu:365,435,515)
template <bool EMIT, class Sink>
__device__ __forceinline__ void ag_put(AgEnc &s, int32_t del, uint32_t bit_size, Sink &sink)
{
    if (s.in_run) {
        if (del == 0) {
            s.nz++;
            s.c++;
            if (s.nz >= 65535u) ag_flush_run<EMIT>(s, sink, 0u);
            else if (s.c == s.count) ag_flush_run<EMIT>(s, sink, 1u);
            return;
        }
        ag_flush_run<EMIT>(s, sink, 1u);
    }
    const uint32_t mb = s.mb;
    const uint32_t k = min(bfind_u32((mb >> kQbShift) + 3u), kKb0);
    const uint32_t m = (1u << k) - 1u;
    const uint32_t n = (uint32_t)((del << 1) ^ (del >> 31)) - s.zmode;           // ag_enc.c:287

    // codec/ag_enc.c:151-184 dyn_code_32bit (numBits can never exceed 25 when div < 9, k <= 14)
    if (n < kMaxPrefix * m) {
        const uint32_t div = (k == 1) ? n : __umulhi(n, c_div_magic[k]);
        const uint32_t mod = n - div * m;
        const uint32_t de = (mod == 0);
        const uint32_t len = div + k + 1 - de;
        s.bits += len;
        if (EMIT) sink.put((((1u << div) - 1u) << (len - div)) + mod + 1 - de, len);
    } else {
        s.bits += kMaxPrefix + bit_size;
        if (EMIT) {
            sink.put((1u << kMaxPrefix) - 1u, kMaxPrefix);
            sink.put(bit_size == 32 ? n : (n & ((1u << bit_size) - 1u)), bit_size);
        }
    }
    s.c++;
    uint32_t nmb = kPb0 * (n + s.zmode) + mb - ((kPb0 * mb) >> kQbShift);       // ag_enc.c:314
    if (n > kMeanClamp) nmb = kMeanClamp;
    s.mb = nmb;
    s.zmode = 0;
    if (((nmb << 2) < kQb) && (s.c < s.count)) {                                  // ag_enc.c:324
        s.in_run = 1;
        s.nz = 0;
    }
}

// ---- MSB-first bit readers over global memory (decoder) -------------------------------------------------
// Positions are in bits from the packet's first byte; words past the packet's last byte read as zero.

// Stateless random-access reader: two aligned word loads per peek.  Used where reads are sparse
// (header pre-pass) or naturally coalesced across a warp (shift bytes in the output kernel).
struct BitPeek {
    const uint32_t *base;   // 4-byte aligned address at or before the packet
    uint32_t bias;          // bit offset of the packet's first bit inside base[0]
    uint32_t last_word;     // index of the last word holding packet bytes
    uint32_t tail_mask;     // the packet's bits inside that word (MSB-first view)
    uint32_t pos;
    bool valid;
    __device__ __forceinline__ void start(const uint8_t *packet, uint32_t nbytes)
    {
        const uintptr_t addr = reinterpret_cast<uintptr_t>(packet);
        base = reinterpret_cast<const uint32_t *>(addr & ~(uintptr_t)3);
        bias = (uint32_t)(addr & 3u) * 8u;
        valid = nbytes != 0;
        last_word = valid ? (bias + nbytes * 8u - 1u) >> 5 : 0u;
        const uint32_t tail_bits = valid ? ((bias + nbytes * 8u - 1u) & 31u) + 1u : 32u;     // valid bits of the last word, from its MSB
        tail_mask = tail_bits == 32u ? 0xffffffffu : ~(0xffffffffu >> tail_bits);
        pos = 0;
    }
    // bytes past the packet's last byte read as zero, also inside its last word
    __device__ __forceinline__ uint32_t word(uint32_t i) const
    {
        if (!(valid && i <= last_word)) return 0u;
        const uint32_t w = bswap32(__ldg(base + i));
        return i == last_word ? (w & tail_mask) : w;
    }
    __device__ __forceinline__ uint32_t peek32_at(uint32_t p) const
    {
        const uint32_t abs_bit = bias + p;
        const uint32_t i = abs_bit >> 5;
        return __funnelshift_l(word(i + 1), word(i), abs_bit & 31u);
    }
    __device__ __forceinline__ uint32_t bits_at(uint32_t p, uint32_t nbits) const { return nbits ? peek32_at(p) >> (32u - nbits) : 0u; }
    __device__ __forceinline__ uint32_t get(uint32_t nbits)
    {
        const uint32_t v = bits_at(pos, nbits);
        pos += nbits;
        return v;
    }
};

// Sequential reader for the serial Golomb decode.  The reader is BRANCH-FREE on the per-symbol path, because the
// lanes of a warp cross word boundaries at different symbols and any "refill if needed" branch would be taken by a
// few lanes at almost every symbol (the warp then pays for it every time):
//   * a 64-bit MSB-first window (hi:lo) with `navail` valid bits; consume() is two funnel shifts;
//   * refill() tops the window up with the pre-loaded word `nxt` when navail <= 32 -- selects and one predicated
//     LDS, no branch -- so hi always holds 32 valid bits;
//   * the packet's words stream through a private shared-memory ring (slot s of this lane is
//     ring[s * kRingStride], bank == lane) filled by cp.async (LDGSTS).  The ring is topped up at a WARP-UNIFORM
//     cadence (every kTopUpEvery symbols, top_up()); what one top-up requests is only assumed to have landed TWO
//     top-ups later (wait_group 2), so the global-load latency never sits on the decode's dependency chain (with
//     wait_group 1 the entropy warp still spent a fifth of its time waiting there).
//     One symbol step consumes at most 9 + 32 bits plus a 25-bit run code = 66 bits, so three periods consume at most
//     3 * kTopUpEvery * 66 bits = 50 words, inside the kRingSlots = 64 words requested ahead.
//   * headers and escape samples are sparse reads and go through BitPeek instead; seek() (synchronous) positions
//     this reader at the first bit of a Golomb stream.
constexpr uint32_t kRingSlots = 64;
constexpr uint32_t kTopUpEvery = 8;
#ifndef ALAC_DEC_LANES
#define ALAC_DEC_LANES 32
#endif
constexpr uint32_t kRingStride = ALAC_DEC_LANES;     // lanes per CTA of the kernels that use BitReader
static_assert(3 * kTopUpEvery * 66 + 4 * 32 <= kRingSlots * 32, "ring too small for the top-up cadence");

// 4-byte cp.async with a source size: src_bytes = 0 reads nothing and zero-fills the destination
__device__ __forceinline__ void cp_async_word(uint32_t smem_dst, const uint32_t *gsrc, uint32_t src_bytes)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t smem_addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_addr) : "memory");
    return v;
}
// v = take ? shared[addr] : v, as a predicated load (never a branch)
__device__ __forceinline__ void lds_u32_if(uint32_t &v, uint32_t smem_addr, bool take)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.shared.u32 %0, [%1];\n\t}"
                 : "+r"(v) : "r"(smem_addr), "r"((uint32_t)take) : "memory");
}

struct BitReader {
    const uint32_t *base;   // 4-byte aligned address at or before the packet
    uint32_t bias;          // bit offset of the packet's first bit inside base[0]
    int32_t last_word;      // index of the last word holding packet bytes; -1 for an empty packet
    uint32_t tail_bytes;    // packet bytes inside that word
    uint32_t ring;          // shared-memory address of this lane's ring column
    uint32_t rd;            // index of the word held in nxt (the next one to enter the window)
    uint32_t wr;            // next word index to request
    uint32_t hi, lo;        // window: the next unread bit is the MSB of hi
    uint32_t navail;        // valid bits in hi:lo (>= 32 after refill())
    uint32_t nxt;           // word rd, byte-swapped
    uint32_t pos;           // bits consumed since the start of the packet

    __device__ __forceinline__ uint32_t slot_addr(uint32_t i) const { return ring + (i & (kRingSlots - 1u)) * (kRingStride * 4u); }
    // branch-free: words past the end are not read, their slot is zero-filled by the copy itself
    __device__ __forceinline__ void issue(uint32_t i)
    {
        const bool in = (int32_t)i <= last_word;
        // the last word is copied only up to the packet's last byte: the copy zero-fills the rest
        cp_async_word(slot_addr(i), base + (in ? i : 0u), in ? ((int32_t)i == last_word ? tail_bytes : 4u) : 0u);
    }
    // request everything up to kRingSlots words past the read position (slots of words < rd are free)
    __device__ __forceinline__ void request_ahead()
    {
        const uint32_t limit = rd + kRingSlots;
        while ((int32_t)(limit - wr) > 0) issue(wr++);
        cp_async_commit();
    }
    // asynchronous top-up: what THIS call requests is only usable after the next two
    __device__ __forceinline__ void top_up()
    {
        request_ahead();
        cp_async_wait<2>();
    }
    // synchronous top-up: everything requested so far is in
    __device__ __forceinline__ void prime()
    {
        request_ahead();
        cp_async_wait<0>();
    }
    __device__ __forceinline__ void consume(uint32_t nbits)     // 0..32 (<= navail)
    {
        hi = __funnelshift_lc(lo, hi, nbits);
        lo = __funnelshift_lc(0u, lo, nbits);
        navail -= nbits;
        pos += nbits;
    }
    __device__ __forceinline__ void refill()
    {
        const bool take = navail <= 32u;
        // insert nxt below the navail valid bits (lo holds none of them when navail <= 32)
        const uint32_t add_hi = __funnelshift_rc(nxt, 0u, navail);      // nxt >> navail, 0 at navail = 32
        const uint32_t new_lo = __funnelshift_rc(0u, nxt, navail);      // low half of (nxt:0) >> navail
        hi |= take ? add_hi : 0u;
        lo = take ? new_lo : lo;
        navail += take ? 32u : 0u;
        rd += take ? 1u : 0u;
        lds_u32_if(nxt, slot_addr(rd), take);
        nxt = take ? bswap32(nxt) : nxt;
    }
    // jump to bit position p of the packet (synchronous)
    __device__ __forceinline__ void seek(uint32_t p)
    {
        const uint32_t abs_bit = bias + p, w = abs_bit >> 5;
        cp_async_wait<0>();         // nothing older may still be writing ring slots
        pos = p;
        rd = w;
        wr = w;
        prime();
        hi = bswap32(lds_u32(slot_addr(w)));
        lo = bswap32(lds_u32(slot_addr(w + 1)));
        nxt = bswap32(lds_u32(slot_addr(w + 2)));
        rd = w + 2;
        navail = 64;
        consume(abs_bit & 31u);
        pos = p;
        refill();
    }
    __device__ __forceinline__ void start(const uint8_t *packet, uint32_t nbytes, uint32_t *ring_column)
    {
        const uintptr_t addr = reinterpret_cast<uintptr_t>(packet);
        base = reinterpret_cast<const uint32_t *>(addr & ~(uintptr_t)3);
        bias = (uint32_t)(addr & 3u) * 8u;
        last_word = nbytes ? (int32_t)((bias + nbytes * 8u - 1u) >> 5) : -1;
        tail_bytes = (((bias >> 3) + nbytes - 1u) & 3u) + 1u;
        if (!nbytes) base = ring_column;        // never dereferenced (every request has size 0), but keep it sane
        ring = (uint32_t)__cvta_generic_to_shared(ring_column);
        pos = 0;
    }
    __device__ __forceinline__ uint32_t peek32() const { return hi; }
};

// streaming dyn_decomp (codec/ag_dec.c:272-362): next() yields one residual per call.  The common symbol
// (codec/ag_dec.c:220-270 dyn_get_32bit, prefix < 9) is straight-line code; the escape code and the zero-run
// code (codec/ag_dec.c:171-217 dyn_get) are the only branches.
struct AgDec {
    uint32_t mb, zmode, count;
    uint32_t next_real;     // index of the next sample that has a code of its own (samples before it are a zero run)
    uint32_t pb, kb, wb, max_size;
    uint32_t start_rel;     // first bit's byte-floor, for the reference's "bitPos < maxPos" test
    uint32_t last_start;    // bit position at which the latest code started
    int32_t status;
    __device__ __forceinline__ void start(const BitReader &br, uint32_t n, uint32_t mb0, uint32_t pb_, uint32_t kb_, uint32_t max_size_)
    {
        mb = mb0; zmode = 0; next_real = 0; count = n;
        pb = pb_; kb = kb_; wb = (1u << kb_) - 1u; max_size = min(max_size_, 32u);
        start_rel = br.pos & ~7u;
        last_start = start_rel;
        status = 0;
    }
    // residual of sample j; call with j = 0, 1, 2, ... (every lane of a warp at the same j: zero runs are
    // waited out, not skipped, so the ring top-up cadence stays warp-uniform).  j >= count yields 0.
    __device__ __forceinline__ int32_t at(BitReader &br, uint32_t j)
    {
        if ((j - next_real) >= (count - next_real)) return 0;       // inside a zero run, or past the end
        last_start = br.pos;
        const uint32_t k = min(bfind_u32((mb >> kQbShift) + 3u), kb);
        const uint32_t m = (1u << k) - 1u;
        const uint32_t window = br.peek32();
        const uint32_t pre = (uint32_t)__clz((int)~window);
        uint32_t n;
        if (pre >= kMaxPrefix) {
            br.consume(kMaxPrefix);
            br.refill();
            n = br.peek32() >> (32u - max_size);
            br.consume(max_size);
        } else {
            // k low bits after the prefix's terminating zero; v < 2 means the code was one bit shorter
            const uint32_t v = (window << (pre + 1u)) >> (32u - k);
            const uint32_t big = (v >= 2u) ? 1u : 0u;
            n = pre * m + (big ? v - 1u : 0u);
            br.consume(pre + k + big);
        }
        br.refill();
        const uint32_t nd = n + zmode;
        const int32_t del = (int32_t)(nd >> 1) ^ -(int32_t)(nd & 1u);                    // ag_dec.c:313-319 (zig-zag)
        next_real = j + 1u;
        mb = pb * nd + mb - ((pb * mb) >> kQbShift);
        if (n > kMeanClamp) mb = kMeanClamp;
        zmode = 0;
        if (((mb << 2) < kQb) && (j + 1u < count)) {                                     // ag_dec.c:334
            zmode = 1;
            const uint32_t kz = (uint32_t)__clz((int)mb) - 24u + ((mb + 16u) >> 6);
            const uint32_t mz = ((1u << kz) - 1u) & wb;
            const uint32_t w2 = br.peek32();
            const uint32_t pz = (uint32_t)__clz((int)~w2);
            uint32_t run;
            if (pz >= kMaxPrefix) {
                run = (w2 << kMaxPrefix) >> (32u - kRunRawBits);
                br.consume(kMaxPrefix + kRunRawBits);
            } else {
                const uint32_t v = (w2 << (pz + 1u)) >> (32u - kz);
                uint32_t nb = pz + 1u + kz;
                run = pz * mz + v - 1u;
                if (v < 2u) { run -= (v - 1u); nb -= 1u; }
                br.consume(nb);
            }
            br.refill();
            if (!(j + 1u + run <= count)) { status = -50; }                              // ag_dec.c:341
            else next_real = j + 1u + run;
            if (run >= 65535u) zmode = 0;
            mb = 0;
        }
        return del;
    }
    // ag_dec.c:302 "bitPos < maxPos" is tested before every code; positions only grow, so testing the start of the
    // LAST code is the same test.  (Decoding runs on over zero-filled words after the end of a packet: a failed
    // packet's samples are unspecified and the hot loop carries no early-out.)  Then dyn_decomp's exit check
    // "cur <= end" (ag_dec.c:359).  cap_bits = packet bytes * 8.
    __device__ __forceinline__ int32_t finish(const BitReader &br, uint32_t cap_bits)
    {
        if (count && !((last_start - start_rel) < cap_bits)) status = -50;
        if (!status && (br.pos >> 3) > (cap_bits >> 3)) status = -50;
        return status;
    }
};

// ---- named barriers: FULL / EMPTY hand-off of shared-memory tiles between the two warps of a 64-thread CTA ----------
enum : uint32_t { BAR_FULL0 = 1, BAR_EMPTY0 = 3 };      // + buffer index

// Barrier numbers are immediates (a register operand makes ptxas reserve all 16 barriers for the CTA, which caps
// the SM at 4 resident CTAs); `second` selects buffer 1.
template <uint32_t ID> __device__ __forceinline__ void bar_sync64() { asm volatile("bar.sync %0, 64;" ::"n"(ID) : "memory"); }
template <uint32_t ID> __device__ __forceinline__ void bar_arrive64() { asm volatile("bar.arrive %0, 64;" ::"n"(ID) : "memory"); }
template <uint32_t ID0> __device__ __forceinline__ void named_sync(bool second)
{
    if (second) bar_sync64<ID0 + 1>(); else bar_sync64<ID0>();
}
template <uint32_t ID0> __device__ __forceinline__ void named_arrive(bool second)
{
    __threadfence_block();      // what this warp wrote to the buffer is visible to the warp that waits
    if (second) bar_arrive64<ID0 + 1>(); else bar_arrive64<ID0>();
}


}  // namespace alacb
