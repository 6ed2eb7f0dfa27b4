/*
 * alac_oracle.c -- CPU oracle (restatement) of the reference ALAC hot path.
 *
 * TEST INFRASTRUCTURE ONLY -- see alac_oracle.h.  Never linked into, imported by
 * or executed from the product path (alac_b200/).
 *
 * Every function cites the reference file:line it follows (relative to
 * /root/reference).  The fork's GPU plumbing (cudaMemcpy of pre-mixed slices,
 * device scratch) is replaced by computing the same slices on the host; the
 * fork's defects listed in SURVEY.md A.4 are NOT replicated (Apple semantics).
 *
 * Deterministic-padding rule (DESIGN.md "tiny tails"): samples of a mix buffer
 * beyond the count that was mixed read as zero.  The reference reads stale
 * (previous-frame / uninitialised device) data there; it only matters for tail
 * frames shorter than 72 samples.
 *
 * Build with -fwrapv -fno-strict-aliasing (int32 MACs wrap, codec/dp_enc.c:228).
 */
#include "alac_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* constants                                                                  */
/* ------------------------------------------------------------------------- */
/* codec/aglib.h:36-55 */
#define QBSHIFT            9
#define QB                 (1u << QBSHIFT)
#define PB0                40
#define MB0                10
#define KB0                14
#define MAX_RUN_DEFAULT    255
#define MMULSHIFT          2
#define MDENSHIFT          (QBSHIFT - MMULSHIFT - 1)
#define MOFF               (1u << (MDENSHIFT - 2))
#define BITOFF             24
#define MAX_PREFIX         9
#define RUN_RAW_BITS       16
#define N_MAX_MEAN_CLAMP   0xffffu
/* codec/dplib.h:40-45 */
#define DENSHIFT_DEFAULT   9
/* codec/ALACEncoder.cu:56-61 */
#define DEFAULT_MIX_BITS   2
#define MAX_RES            4
#define DEFAULT_NUM_UV     8
#define MIN_UV             4
#define MAX_UV             8
/* codec/ALACAudioTypes.h:68-75 */
#define MAX_CHANNELS       8
#define MAX_SEARCHES       16
#define MAX_COEFS          16

/* ------------------------------------------------------------------------- */
/* bit cursor (codec/ALACBitUtilities.c:42-62, 212-249: MSB-first)            */
/* ------------------------------------------------------------------------- */
void orc_bits_init(orc_bits *b, uint8_t *buf, uint64_t cap_bytes)
{
    b->buf = buf;
    b->pos = 0;
    b->cap = cap_bytes * 8;
}

void orc_put(orc_bits *b, uint32_t value, unsigned nbits)
{
    while (nbits) {
        unsigned room = 8u - (unsigned)(b->pos & 7u);
        unsigned take = nbits < room ? nbits : room;
        unsigned down = room - take;
        uint32_t piece = (value >> (nbits - take)) & ((1u << take) - 1u);
        uint8_t  mask = (uint8_t)(((1u << take) - 1u) << down);
        uint8_t *p = b->buf + (b->pos >> 3);
        *p = (uint8_t)((*p & ~mask) | (uint8_t)(piece << down));
        b->pos += take;
        nbits -= take;
    }
}

static inline uint64_t load_be64(const uint8_t *p)
{
    uint64_t v = 0;
    for (int i = 0; i < 8; i++) v = (v << 8) | p[i];
    return v;
}

/* extract nbits (0..32) at absolute bit position pos; buffer must have 8 spare bytes */
static inline uint32_t peek_bits(const uint8_t *buf, uint64_t pos, unsigned nbits)
{
    if (nbits == 0) return 0;
    uint64_t w = load_be64(buf + (pos >> 3)) << (pos & 7u);
    return (uint32_t)(w >> (64 - nbits));
}

uint32_t orc_get(orc_bits *b, unsigned nbits)
{
    uint32_t v = peek_bits(b->buf, b->pos, nbits);
    b->pos += nbits;
    return v;
}

/* ------------------------------------------------------------------------- */
/* small helpers                                                              */
/* ------------------------------------------------------------------------- */
/* codec/dp_enc.c:69-75 sign_of_int */
static inline int32_t sign3(int32_t v) { return (v > 0) - (v < 0); }

/* the "(x << chanshift) >> chanshift" idiom, codec/dp_enc.c:232 */
static inline int32_t wrap_bits(int32_t v, uint32_t chanbits)
{
    uint32_t sh = 32u - chanbits;
    return (int32_t)((uint32_t)v << sh) >> sh;
}

/* codec/ag_enc.c:65-79 lead(): leading zeros of a 32-bit word, 32 for zero */
static inline uint32_t lead_zeros(uint32_t v) { return v ? (uint32_t)__builtin_clz(v) : 32u; }

/* codec/ag_enc.c:83-91 lg3a */
static inline uint32_t lg3a(uint32_t m) { return 31u - lead_zeros(m + 3u); }

uint32_t orc_fnv1a(const void *data, size_t n)
{
    const uint8_t *p = (const uint8_t *)data;
    uint32_t h = 2166136261u;
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 16777619u; }
    return h;
}

/* ------------------------------------------------------------------------- */
/* dynamic predictor                                                          */
/* ------------------------------------------------------------------------- */
/* codec/dp_enc.c:49-60 */
void orc_init_coefs(int16_t *coefs, uint32_t denshift, int32_t n)
{
    int32_t den = 1 << denshift;
    for (int32_t k = 0; k < n; k++) coefs[k] = 0;
    coefs[0] = (int16_t)((38 * den) >> 4);
    coefs[1] = (int16_t)((-29 * den) >> 4);
    coefs[2] = (int16_t)((-2 * den) >> 4);
}

/* sign-LMS walk shared by encode and decode: taps visited last-to-first,
   codec/dp_enc.c:236-329 (unrolled) == :362-385 (general); codec/dp_dec.c:358-379 */
static inline void lms_adapt(int16_t *coefs, const int32_t *newest, int32_t top,
                             int32_t numactive, int32_t err, uint32_t denshift)
{
    int32_t left = err;
    if (err > 0) {
        for (int32_t k = numactive - 1; k >= 0; k--) {
            int32_t dd = (int32_t)((uint32_t)top - (uint32_t)newest[-k]);
            int32_t s = sign3(dd);
            coefs[k] = (int16_t)(coefs[k] - s);
            left -= (numactive - k) * ((int32_t)((uint32_t)s * (uint32_t)dd) >> denshift);
            if (left <= 0) break;
        }
    } else if (err < 0) {
        for (int32_t k = numactive - 1; k >= 0; k--) {
            int32_t dd = (int32_t)((uint32_t)top - (uint32_t)newest[-k]);
            int32_t s = sign3(dd);
            coefs[k] = (int16_t)(coefs[k] + s);
            /* arithmetic shift of a non-positive product: rounds toward -inf (dp_enc.c:288) */
            left -= (numactive - k) * ((int32_t)((uint32_t)(-s) * (uint32_t)dd) >> denshift);
            if (left >= 0) break;
        }
    }
}

/* codec/dp_enc.c:77-388 pc_block */
void orc_pc_block(const int32_t *in, int32_t *res, int32_t num, int16_t *coefs,
                  int32_t numactive, uint32_t chanbits, uint32_t denshift)
{
    res[0] = in[0];
    if (numactive == 0) {                       /* :91-97 copy mode */
        if (num > 1 && in != res) memmove(&res[1], &in[1], (size_t)(num - 1) * sizeof(int32_t));
        return;
    }
    if (numactive == 31) {                      /* :98-107 first-difference mode */
        for (int32_t j = 1; j < num; j++)
            res[j] = wrap_bits((int32_t)((uint32_t)in[j] - (uint32_t)in[j - 1]), chanbits);
        return;
    }
    /* :108-112 warm-up runs over numactive samples regardless of num */
    for (int32_t j = 1; j <= numactive; j++)
        res[j] = wrap_bits((int32_t)((uint32_t)in[j] - (uint32_t)in[j - 1]), chanbits);

    const int32_t half = 1 << (denshift - 1);
    for (int32_t j = numactive + 1; j < num; j++) {
        const int32_t top = in[j - numactive - 1];
        const int32_t *newest = in + j - 1;
        uint32_t acc = 0;
        for (int32_t k = 0; k < numactive; k++)
            acc += (uint32_t)(int32_t)coefs[k] * ((uint32_t)newest[-k] - (uint32_t)top);
        int32_t pred = (int32_t)(acc + (uint32_t)half) >> denshift;
        int32_t err = wrap_bits((int32_t)((uint32_t)in[j] - (uint32_t)top - (uint32_t)pred), chanbits);
        res[j] = err;
        lms_adapt(coefs, newest, top, numactive, err, denshift);
    }
}

/* codec/dp_dec.c:55-381 unpc_block */
void orc_unpc_block(const int32_t *res, int32_t *out, int32_t num, int16_t *coefs,
                    int32_t numactive, uint32_t chanbits, uint32_t denshift)
{
    out[0] = res[0];
    if (numactive == 0) {                       /* :67-73 */
        if (num > 1 && res != out) memmove(&out[1], &res[1], (size_t)(num - 1) * sizeof(int32_t));
        return;
    }
    if (numactive == 31) {                      /* :74-95 */
        int32_t prev = out[0];
        for (int32_t j = 1; j < num; j++) {
            prev = wrap_bits((int32_t)((uint32_t)res[j] + (uint32_t)prev), chanbits);
            out[j] = prev;
        }
        return;
    }
    for (int32_t j = 1; j <= numactive; j++)    /* :97-101 */
        out[j] = wrap_bits((int32_t)((uint32_t)res[j] + (uint32_t)out[j - 1]), chanbits);

    const int32_t half = 1 << (denshift - 1);
    for (int32_t j = numactive + 1; j < num; j++) {
        const int32_t top = out[j - numactive - 1];
        const int32_t *newest = out + j - 1;
        uint32_t acc = 0;
        for (int32_t k = 0; k < numactive; k++)
            acc += (uint32_t)(int32_t)coefs[k] * ((uint32_t)newest[-k] - (uint32_t)top);
        int32_t err = res[j];
        int32_t pred = (int32_t)(acc + (uint32_t)half) >> denshift;
        out[j] = wrap_bits((int32_t)((uint32_t)err + (uint32_t)top + (uint32_t)pred), chanbits);
        lms_adapt(coefs, newest, top, numactive, err, denshift);
    }
}

/* ------------------------------------------------------------------------- */
/* adaptive Golomb coder                                                      */
/* ------------------------------------------------------------------------- */
/* codec/ag_dec.c:73-83 (only the fields the coder reads) */
void orc_ag_params_set(orc_ag_params *p, uint32_t mb0, uint32_t pb, uint32_t kb)
{
    p->mb0 = mb0;
    p->pb = pb;
    p->kb = kb;
    p->wb = (1u << kb) - 1u;
}

/* codec/ag_enc.c:151-184 dyn_code_32bit + the two jams at :289-300 */
static inline void put_sample_code(orc_bits *b, uint32_t m, uint32_t k, uint32_t n, int32_t bit_size)
{
    uint32_t div = n / m;
    if (div < MAX_PREFIX) {
        uint32_t mod = n - m * div;
        uint32_t de = (mod == 0);
        uint32_t len = div + k + 1 - de;
        if (len <= 25) {
            uint32_t value = (((1u << div) - 1u) << (len - div)) + mod + 1 - de;
            orc_put(b, value, len);
            return;
        }
    }
    orc_put(b, (1u << MAX_PREFIX) - 1u, MAX_PREFIX);
    orc_put(b, n, (unsigned)bit_size);
}

/* codec/ag_enc.c:115-148 dyn_code (zero-run lengths, 16-bit escape payload) */
static inline void put_run_code(orc_bits *b, uint32_t m, uint32_t k, uint32_t n)
{
    uint32_t div = n / m;
    if (div < MAX_PREFIX) {
        uint32_t mod = n % m;
        uint32_t de = (mod == 0);
        uint32_t len = div + k + 1 - de;
        if (len <= MAX_PREFIX + RUN_RAW_BITS) {
            uint32_t value = (((1u << div) - 1u) << (len - div)) + mod + 1 - de;
            orc_put(b, value, len);
            return;
        }
    }
    orc_put(b, ((((1u << MAX_PREFIX) - 1u)) << RUN_RAW_BITS) + n, MAX_PREFIX + RUN_RAW_BITS);
}

/* codec/ag_enc.c:249-367 dyn_comp */
int32_t orc_dyn_comp(const orc_ag_params *p, const int32_t *res, orc_bits *b,
                     int32_t num, int32_t bit_size, uint32_t *out_bits)
{
    *out_bits = 0;
    if (bit_size < 1 || bit_size > 32) return ORC_PARAM_ERROR;

    const uint64_t start = b->pos;
    const uint32_t pb = p->pb, kb = p->kb, wb = p->wb;
    uint32_t mb = p->mb0;
    uint32_t zmode = 0;
    int32_t c = 0;

    while (c < num) {
        uint32_t k = lg3a(mb >> QBSHIFT);
        if (k > kb) k = kb;
        uint32_t m = (1u << k) - 1u;

        int32_t del = res[c++];
        uint32_t mag = (uint32_t)(del < 0 ? -(uint32_t)del : (uint32_t)del);
        uint32_t n = (mag << 1) - ((uint32_t)del >> 31) - zmode;       /* :287 */

        put_sample_code(b, m, k, n, bit_size);

        mb = pb * (n + zmode) + mb - ((pb * mb) >> QBSHIFT);            /* :314 */
        if (n > N_MAX_MEAN_CLAMP) mb = N_MAX_MEAN_CLAMP;                /* :317-318 */
        zmode = 0;

        if (((mb << MMULSHIFT) < QB) && (c < num)) {                    /* :324-361 */
            uint32_t nz = 0;
            zmode = 1;
            while (c < num && res[c] == 0) {
                c++;
                nz++;
                if (nz >= 65535) { zmode = 0; break; }
            }
            k = lead_zeros(mb) - BITOFF + ((mb + MOFF) >> MDENSHIFT);
            uint32_t mz = ((1u << k) - 1u) & wb;
            put_run_code(b, mz, k, nz);
            mb = 0;
        }
    }
    *out_bits = (uint32_t)(b->pos - start);
    return ORC_OK;
}

/* codec/ag_dec.c:220-270 dyn_get_32bit */
static inline uint32_t get_sample_code(const uint8_t *buf, uint64_t *pos, uint32_t m, uint32_t k, int32_t maxbits)
{
    uint64_t at = *pos;
    uint32_t window = peek_bits(buf, at, 32);
    uint32_t pre = lead_zeros(~window);
    uint32_t result;

    if (pre >= MAX_PREFIX) {
        result = peek_bits(buf, at + MAX_PREFIX, (unsigned)maxbits);
        at += MAX_PREFIX + (uint32_t)maxbits;
    } else {
        result = pre;
        at += pre + 1;
        if (k != 1) {
            uint32_t v = (window << (pre + 1)) >> (32 - k);
            at += k - 1;
            result = pre * m;
            if (v >= 2) { result += v - 1; at += 1; }
        }
    }
    *pos = at;
    return result;
}

/* codec/ag_dec.c:171-217 dyn_get */
static inline uint32_t get_run_code(const uint8_t *buf, uint64_t *pos, uint32_t m, uint32_t k)
{
    uint64_t at = *pos;
    uint32_t window = peek_bits(buf, at, 32);
    uint32_t pre = lead_zeros(~window);
    uint32_t result;

    if (pre >= MAX_PREFIX) {
        result = (window << MAX_PREFIX) >> (32 - RUN_RAW_BITS);
        at += MAX_PREFIX + RUN_RAW_BITS;
    } else {
        uint32_t v = (window << (pre + 1)) >> (32 - k);
        at += pre + 1 + k;
        result = pre * m + v - 1;
        if (v < 2) { result -= (v - 1); at -= 1; }
    }
    *pos = at;
    return result;
}

/* codec/ag_dec.c:272-362 dyn_decomp.  b->cap plays bitstream->byteSize*8; the reference's
   bitPos counts from the byte the cursor sits in (`in = bitstream->cur`, :288-291). */
int32_t orc_dyn_decomp(const orc_ag_params *p, orc_bits *b, int32_t *res,
                       int32_t num, int32_t max_size, uint32_t *out_bits)
{
    *out_bits = 0;
    const uint8_t *buf = b->buf;
    const uint64_t start = b->pos;
    const uint64_t rel0 = start & ~(uint64_t)7;
    const uint32_t pb = p->pb, kb = p->kb, wb = p->wb;
    uint64_t pos = start;
    uint32_t mb = p->mb0;
    uint32_t zmode = 0;
    int32_t c = 0;
    int32_t status = ORC_OK;

    while (c < num) {
        if (!((pos - rel0) < b->cap)) { status = ORC_PARAM_ERROR; break; }       /* :302 */

        uint32_t k = lg3a(mb >> QBSHIFT);
        if (k > kb) k = kb;
        uint32_t m = (1u << k) - 1u;

        uint32_t n = get_sample_code(buf, &pos, m, k, max_size);

        uint32_t nd = n + zmode;                                                  /* :313-319 */
        int32_t mult = -(int32_t)(nd & 1u);
        mult |= 1;
        res[c++] = (int32_t)(((nd + 1u) >> 1) * (uint32_t)mult);

        mb = pb * (n + zmode) + mb - ((pb * mb) >> QBSHIFT);
        if (n > N_MAX_MEAN_CLAMP) mb = N_MAX_MEAN_CLAMP;
        zmode = 0;

        if (((mb << MMULSHIFT) < QB) && (c < num)) {                              /* :334-357 */
            zmode = 1;
            k = lead_zeros(mb) - BITOFF + ((mb + MOFF) >> MDENSHIFT);
            uint32_t mz = ((1u << k) - 1u) & wb;
            n = get_run_code(buf, &pos, mz, k);
            if (!((uint64_t)c + n <= (uint64_t)num)) { status = ORC_PARAM_ERROR; break; }   /* :341 */
            for (uint32_t j = 0; j < n; j++) res[c++] = 0;
            if (n >= 65535) zmode = 0;
            mb = 0;
        }
    }
    *out_bits = (uint32_t)(pos - start);
    b->pos = pos;
    if ((b->pos >> 3) > (b->cap >> 3)) status = ORC_PARAM_ERROR;                  /* :359 cur <= end */
    return status;
}

static const orc_prims g_port_prims = {
    "port", orc_pc_block, orc_unpc_block, orc_dyn_comp, orc_dyn_decomp
};
const orc_prims *orc_prims_port(void) { return &g_port_prims; }

#ifndef ORC_WITH_REFERENCE
const orc_prims *orc_prims_reference(void) { return NULL; }
#endif

static const orc_prims *pick_prims(int flavour)
{
    return flavour ? orc_prims_reference() : orc_prims_port();
}

/* ------------------------------------------------------------------------- */
/* PCM access: packed little-endian, interleaved                               */
/* ------------------------------------------------------------------------- */
static inline uint32_t bytes_per_sample(uint32_t depth) { return depth == 16 ? 2u : depth == 32 ? 4u : 3u; }

/* full-width sample, right-aligned and sign-extended.
   16: codec/matrix_enc.cu:72-99; 20: :120-159 "(l<<8)>>12"; 24: :186-213 "(l<<8)>>8"; 32: :330-353 */
static inline int32_t pcm_fetch(const uint8_t *pcm, uint32_t depth, uint64_t idx)
{
    const uint8_t *p;
    switch (depth) {
    case 16:
        p = pcm + idx * 2;
        return (int16_t)((uint16_t)p[0] | ((uint16_t)p[1] << 8));
    case 20:
        p = pcm + idx * 3;
        return (int32_t)((((uint32_t)p[2] << 16) | ((uint32_t)p[1] << 8) | p[0]) << 8) >> 12;
    case 24:
        p = pcm + idx * 3;
        return (int32_t)((((uint32_t)p[2] << 16) | ((uint32_t)p[1] << 8) | p[0]) << 8) >> 8;
    default:
        p = pcm + idx * 4;
        return (int32_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
    }
}

/* ------------------------------------------------------------------------- */
/* encoder                                                                    */
/* ------------------------------------------------------------------------- */
struct orc_encoder {
    uint32_t channels, bit_depth, sample_rate, frame_size;
    int      fast_mode;
    const orc_prims *prims;
    /* codec/ALACEncoder.h:89-90 */
    int16_t coefs_u[MAX_CHANNELS][MAX_SEARCHES][MAX_COEFS];
    int16_t coefs_v[MAX_CHANNELS][MAX_SEARCHES][MAX_COEFS];
    int32_t *mix_u, *mix_v, *pred_u, *pred_v;
    uint16_t *shift_uv;
    uint8_t *work;
    uint32_t work_bytes;
    uint32_t total_bytes, max_frame_bytes, avg_bit_rate;
};

/* codec/ALACEncoder.cu:97-107 sChannelMaps */
static const uint32_t g_channel_maps[MAX_CHANNELS] = {
    ORC_ID_SCE,
    ORC_ID_CPE,
    (ORC_ID_CPE << 3) | (ORC_ID_SCE),
    (ORC_ID_SCE << 9) | (ORC_ID_CPE << 3) | (ORC_ID_SCE),
    (ORC_ID_CPE << 9) | (ORC_ID_CPE << 3) | (ORC_ID_SCE),
    (ORC_ID_SCE << 15) | (ORC_ID_CPE << 9) | (ORC_ID_CPE << 3) | (ORC_ID_SCE),
    (ORC_ID_SCE << 18) | (ORC_ID_SCE << 15) | (ORC_ID_CPE << 9) | (ORC_ID_CPE << 3) | (ORC_ID_SCE),
    (ORC_ID_SCE << 21) | (ORC_ID_CPE << 15) | (ORC_ID_CPE << 9) | (ORC_ID_CPE << 3) | (ORC_ID_SCE)
};

/* codec/ALACAudioTypes.h:115-125 */
static const uint32_t g_layout_tags[MAX_CHANNELS] = {
    (100u << 16) | 1, (101u << 16) | 2, (113u << 16) | 3, (116u << 16) | 4,
    (120u << 16) | 5, (124u << 16) | 6, (142u << 16) | 7, (127u << 16) | 8
};

/* codec/ALACEncoder.cu:1524-1531 */
void orc_encoder_reset(orc_encoder *e)
{
    for (uint32_t ch = 0; ch < MAX_CHANNELS; ch++)
        for (uint32_t s = 0; s < MAX_SEARCHES; s++) {
            orc_init_coefs(e->coefs_u[ch][s], DENSHIFT_DEFAULT, MAX_COEFS);
            orc_init_coefs(e->coefs_v[ch][s], DENSHIFT_DEFAULT, MAX_COEFS);
        }
}

/* codec/ALACEncoder.cu:1457-1535 InitializeEncoder */
orc_encoder *orc_encoder_new(uint32_t channels, uint32_t bit_depth, uint32_t sample_rate,
                             uint32_t frame_size, int fast_mode, int flavour)
{
    if (channels < 1 || channels > MAX_CHANNELS) return NULL;
    if (!(bit_depth == 16 || bit_depth == 20 || bit_depth == 24 || bit_depth == 32)) return NULL;
    if (frame_size < 1) return NULL;
    const orc_prims *prims = pick_prims(flavour);
    if (!prims) return NULL;

    orc_encoder *e = (orc_encoder *)calloc(1, sizeof(*e));
    if (!e) return NULL;
    e->channels = channels;
    e->bit_depth = bit_depth;
    e->sample_rate = sample_rate;
    e->frame_size = frame_size;
    e->fast_mode = fast_mode;
    e->prims = prims;
    size_t n = (size_t)frame_size + 64;            /* pc_block touches >= numactive+1 entries */
    e->mix_u = (int32_t *)calloc(n, sizeof(int32_t));
    e->mix_v = (int32_t *)calloc(n, sizeof(int32_t));
    e->pred_u = (int32_t *)calloc(n, sizeof(int32_t));
    e->pred_v = (int32_t *)calloc(n, sizeof(int32_t));
    e->shift_uv = (uint16_t *)calloc(2 * n, sizeof(uint16_t));
    e->work_bytes = frame_size * 2u * 5u + 64u;    /* :1489 worst case for one pair, + jam slack */
    e->work = (uint8_t *)calloc(e->work_bytes, 1);
    if (!e->mix_u || !e->mix_v || !e->pred_u || !e->pred_v || !e->shift_uv || !e->work) {
        orc_encoder_free(e);
        return NULL;
    }
    orc_encoder_reset(e);
    return e;
}

void orc_encoder_free(orc_encoder *e)
{
    if (!e) return;
    free(e->mix_u); free(e->mix_v); free(e->pred_u); free(e->pred_v);
    free(e->shift_uv); free(e->work);
    free(e);
}

void orc_encoder_get_coefs(const orc_encoder *e, uint32_t channel, int is_v, uint32_t row, int16_t *out16)
{
    memcpy(out16, is_v ? e->coefs_v[channel][row] : e->coefs_u[channel][row], MAX_COEFS * sizeof(int16_t));
}

static inline void put_be32(uint8_t *p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }
static inline void put_be16(uint8_t *p, uint16_t v) { p[0] = v >> 8; p[1] = (uint8_t)v; }

/* codec/ALACEncoder.cu:1082-1140 GetConfig + GetMagicCookie; layout codec/ALACAudioTypes.h:162-176 */
uint32_t orc_encoder_cookie(const orc_encoder *e, uint8_t *out, uint32_t cap)
{
    uint32_t size = 24 + (e->channels > 2 ? 24u : 0u);
    if (cap < size) return 0;                          /* :1136-1139 "no incomplete cookies" */
    memset(out, 0, size);
    put_be32(out + 0, e->frame_size);
    out[4] = 0;                                        /* compatibleVersion */
    out[5] = (uint8_t)e->bit_depth;
    out[6] = PB0;
    out[7] = MB0;
    out[8] = KB0;
    out[9] = (uint8_t)e->channels;
    put_be16(out + 10, MAX_RUN_DEFAULT);
    put_be32(out + 12, e->max_frame_bytes);
    put_be32(out + 16, e->avg_bit_rate);
    put_be32(out + 20, e->sample_rate);
    if (e->channels > 2) {
        static const uint8_t atom[12] = { 0, 0, 0, 24, 'c', 'h', 'a', 'n', 0, 0, 0, 0 };
        memcpy(out + 24, atom, 12);
        /* :1120 the layout tag is stored native-endian (not swapped): little-endian host */
        uint32_t tag = g_layout_tags[e->channels - 1];
        out[36] = (uint8_t)tag; out[37] = (uint8_t)(tag >> 8); out[38] = (uint8_t)(tag >> 16); out[39] = (uint8_t)(tag >> 24);
    }
    return size;
}

static inline uint32_t shift_bytes_for(uint32_t depth)
{
    /* codec/ALACEncoder.cu:327-332 */
    return depth == 32 ? 2u : depth >= 24 ? 1u : 0u;
}

/* stereo matrixing of `count` sample-frames; zero beyond (deterministic-padding rule).
   codec/matrix_enc.cu:72-99 (16), :120-159 (20), :186-282 (24), :330-391 (32);
   search slices codec/ALACEncoder.cu:1144-1310 */
static void mix_pair(const orc_encoder *e, const uint8_t *pcm, uint32_t stride, uint32_t count,
                     int32_t mix_res, uint32_t shift, int32_t *u, int32_t *v, uint16_t *shift_uv,
                     uint32_t clear_to)
{
    const uint32_t mask = (1u << shift) - 1u;
    const int32_t m2 = (1 << DEFAULT_MIX_BITS) - mix_res;
    for (uint32_t i = 0; i < count; i++) {
        int32_t l = pcm_fetch(pcm, e->bit_depth, (uint64_t)i * stride);
        int32_t r = pcm_fetch(pcm, e->bit_depth, (uint64_t)i * stride + 1);
        if (shift) {
            if (shift_uv) {
                shift_uv[2 * i + 0] = (uint16_t)((uint32_t)l & mask);
                shift_uv[2 * i + 1] = (uint16_t)((uint32_t)r & mask);
            }
            l >>= shift;
            r >>= shift;
        }
        if (mix_res != 0) {
            u[i] = (mix_res * l + m2 * r) >> DEFAULT_MIX_BITS;
            v[i] = l - r;
        } else {
            u[i] = l;
            v[i] = r;
        }
    }
    for (uint32_t i = count; i < count + clear_to; i++) { u[i] = 0; v[i] = 0; }
}

/* codec/ALACEncoder.cu:1312-1382 copyNNToPredictor (+ shift split) */
static void copy_mono(const orc_encoder *e, const uint8_t *pcm, uint32_t stride, uint32_t count,
                      uint32_t shift, int32_t *u, uint16_t *shift_u, uint32_t clear_to)
{
    const uint32_t mask = (1u << shift) - 1u;
    for (uint32_t i = 0; i < count; i++) {
        int32_t s = pcm_fetch(pcm, e->bit_depth, (uint64_t)i * stride);
        if (shift) {
            shift_u[i] = (uint16_t)((uint32_t)s & mask);
            s >>= shift;
        }
        u[i] = s;
    }
    for (uint32_t i = count; i < count + clear_to; i++) u[i] = 0;
}

static uint32_t trial_bits(orc_encoder *e, const int32_t *res, uint32_t count, uint32_t chan_bits)
{
    orc_ag_params ag;
    orc_bits work;
    uint32_t bits = 0;
    orc_bits_init(&work, e->work, e->work_bytes);
    orc_ag_params_set(&ag, MB0, (4 * PB0) / 4, KB0);
    e->prims->golomb_enc(&ag, res, &work, (int32_t)count, (int32_t)chan_bits, &bits);
    return bits;
}

/* codec/ALACEncoder.cu:749-806 EncodeStereoEscape */
static void write_stereo_escape(const orc_encoder *e, orc_bits *b, const uint8_t *pcm, uint32_t stride, uint32_t n)
{
    uint32_t partial = (n == e->frame_size) ? 0u : 1u;
    orc_put(b, 0, 12);
    orc_put(b, (partial << 3) | 1u, 4);
    if (partial) orc_put(b, n, 32);
    for (uint32_t i = 0; i < n; i++) {
        orc_put(b, (uint32_t)pcm_fetch(pcm, e->bit_depth, (uint64_t)i * stride), e->bit_depth);
        orc_put(b, (uint32_t)pcm_fetch(pcm, e->bit_depth, (uint64_t)i * stride + 1), e->bit_depth);
    }
}

/* mono escape: missing from the fork (codec/ALACEncoder.cu:959-963 falls through);
   syntax taken from the decoder's SCE escape branch codec/ALACDecoder.cu:697-727 */
static void write_mono_escape(const orc_encoder *e, orc_bits *b, const uint8_t *pcm, uint32_t stride, uint32_t n)
{
    uint32_t partial = (n == e->frame_size) ? 0u : 1u;
    orc_put(b, 0, 12);
    orc_put(b, (partial << 3) | 1u, 4);
    if (partial) orc_put(b, n, 32);
    for (uint32_t i = 0; i < n; i++)
        orc_put(b, (uint32_t)pcm_fetch(pcm, e->bit_depth, (uint64_t)i * stride), e->bit_depth);
}

static void write_channel_header(orc_bits *b, uint32_t num, const int16_t *coefs)
{
    /* codec/ALACEncoder.cu:477-485: mode 0, denShift 9, pbFactor 4 */
    orc_put(b, (0u << 4) | DENSHIFT_DEFAULT, 8);
    orc_put(b, (4u << 5) | num, 8);
    for (uint32_t i = 0; i < num; i++) orc_put(b, (uint32_t)(int32_t)coefs[i], 16);
}

/* codec/ALACEncoder.cu:290-558 EncodeStereo (fast_mode: :564-743 EncodeStereoFast) */
static int32_t encode_pair(orc_encoder *e, orc_bits *b, const uint8_t *pcm, uint32_t stride,
                           uint32_t channel_index, uint32_t n, orc_trace *tr)
{
    const orc_prims *P = e->prims;
    int16_t (*cu)[MAX_COEFS] = e->coefs_u[channel_index];
    int16_t (*cv)[MAX_COEFS] = e->coefs_v[channel_index];
    const uint32_t bytes_shifted = shift_bytes_for(e->bit_depth);
    const uint32_t shift = bytes_shifted * 8;
    const uint32_t chan_bits = e->bit_depth - shift + 1;                      /* :334 */
    const uint32_t partial = (n == e->frame_size) ? 0u : 1u;                  /* :337 */
    const uint32_t clear_to = 32;   /* entries zeroed past the mixed count */
    const uint64_t start = b->pos;
    uint32_t num_u = DEFAULT_NUM_UV, num_v = DEFAULT_NUM_UV;
    int32_t best_res = 0;
    uint32_t min_bits1, min_bits2, min_bits;
    uint32_t bits1 = 0, bits2 = 0;
    int do_escape;
    orc_ag_params ag;

    const uint32_t escape_bits = (n * e->bit_depth * 2) + (partial ? 32u : 0u) + 16u;   /* :459 */

    if (!e->fast_mode) {
        /* stage A: mixRes search on the first n/8 samples, all on row 7 (:353-379) */
        uint32_t na = n / 8;
        min_bits1 = 1u << 31;
        for (int32_t r = 0; r <= MAX_RES; r++) {
            mix_pair(e, pcm, stride, na, r, shift, e->mix_u, e->mix_v, NULL, clear_to);
            P->predict_enc(e->mix_u, e->pred_u, (int32_t)na, cu[DEFAULT_NUM_UV - 1], DEFAULT_NUM_UV, chan_bits, DENSHIFT_DEFAULT);
            P->predict_enc(e->mix_v, e->pred_v, (int32_t)na, cv[DEFAULT_NUM_UV - 1], DEFAULT_NUM_UV, chan_bits, DENSHIFT_DEFAULT);
            /* one work buffer, U then V appended, as :358-370 */
            orc_bits work;
            orc_bits_init(&work, e->work, e->work_bytes);
            orc_ag_params_set(&ag, MB0, (4 * PB0) / 4, KB0);
            P->golomb_enc(&ag, e->pred_u, &work, (int32_t)na, (int32_t)chan_bits, &bits1);
            P->golomb_enc(&ag, e->pred_v, &work, (int32_t)na, (int32_t)chan_bits, &bits2);
            if (bits1 + bits2 < min_bits1) { min_bits1 = bits1 + bits2; best_res = r; }
        }
        /* full-frame mix with the winner (:385-415) */
        mix_pair(e, pcm, stride, n, best_res, shift, e->mix_u, e->mix_v, e->shift_uv, clear_to);

        /* stage B: numU/numV search (:418-452); the predictor runs over n/32 samples but the
           cost covers n/8, so pred[n/32..n/8) still holds the mixRes=4 trial residuals (F4) */
        num_u = num_v = MIN_UV;
        min_bits1 = min_bits2 = 1u << 31;
        for (uint32_t nuv = MIN_UV; nuv <= MAX_UV; nuv += 4) {
            for (int pass = 0; pass < 8; pass++) {
                P->predict_enc(e->mix_u, e->pred_u, (int32_t)(n / 32), cu[nuv - 1], (int32_t)nuv, chan_bits, DENSHIFT_DEFAULT);
                P->predict_enc(e->mix_v, e->pred_v, (int32_t)(n / 32), cv[nuv - 1], (int32_t)nuv, chan_bits, DENSHIFT_DEFAULT);
            }
            bits1 = trial_bits(e, e->pred_u, n / 8, chan_bits);
            if (bits1 * 8 + 16 * nuv < min_bits1) { min_bits1 = bits1 * 8 + 16 * nuv; num_u = nuv; }
            bits2 = trial_bits(e, e->pred_v, n / 8, chan_bits);
            if (bits2 * 8 + 16 * nuv < min_bits2) { min_bits2 = bits2 * 8 + 16 * nuv; num_v = nuv; }
        }
        /* escape estimate (:455-461) */
        min_bits = min_bits1 + min_bits2 + 64 + (partial ? 32u : 0u);
        if (bytes_shifted) min_bits += n * shift * 2;
        do_escape = (min_bits >= escape_bits);
    } else {
        /* fast mode: mixRes 0, 8 taps, no search; escape decided after the fact (:618-725) */
        mix_pair(e, pcm, stride, n, 0, shift, e->mix_u, e->mix_v, e->shift_uv, clear_to);
        do_escape = 0;
    }

    if (tr) {
        memset(tr, 0, sizeof(*tr));
        tr->tag = ORC_ID_CPE;
        tr->mix_res = best_res;
        tr->num_u = (int32_t)num_u;
        tr->num_v = (int32_t)num_v;
        memcpy(tr->hdr_coefs_u, cu[num_u - 1], 8 * sizeof(int16_t));
        memcpy(tr->hdr_coefs_v, cv[num_v - 1], 8 * sizeof(int16_t));
        tr->escape = do_escape ? 1 : 0;
    }

    if (!do_escape) {
        /* header (:466-485) */
        orc_put(b, 0, 12);
        orc_put(b, (partial << 3) | (bytes_shifted << 1), 4);
        if (partial) orc_put(b, n, 32);
        orc_put(b, DEFAULT_MIX_BITS, 8);
        orc_put(b, (uint32_t)best_res, 8);
        write_channel_header(b, num_u, cu[num_u - 1]);
        write_channel_header(b, num_v, cv[num_v - 1]);
        /* interleaved shift values (:488-500) */
        if (bytes_shifted)
            for (uint32_t i = 0; i < n; i++)
                orc_put(b, ((uint32_t)e->shift_uv[2 * i] << shift) | e->shift_uv[2 * i + 1], shift * 2);
        /* stage C: final pass, U then V (:507-531) */
        P->predict_enc(e->mix_u, e->pred_u, (int32_t)n, cu[num_u - 1], (int32_t)num_u, chan_bits, DENSHIFT_DEFAULT);
        orc_ag_params_set(&ag, MB0, (4 * PB0) / 4, KB0);
        int32_t st = P->golomb_enc(&ag, e->pred_u, b, (int32_t)n, (int32_t)chan_bits, &bits1);
        if (st) return st;
        P->predict_enc(e->mix_v, e->pred_v, (int32_t)n, cv[num_v - 1], (int32_t)num_v, chan_bits, DENSHIFT_DEFAULT);
        st = P->golomb_enc(&ag, e->pred_v, b, (int32_t)n, (int32_t)chan_bits, &bits2);
        if (st) return st;
        if (tr) { tr->bits_u = bits1; tr->bits_v = bits2; }

        if (e->fast_mode) {
            /* :703-713 */
            min_bits = (bits1 + num_u * 16) + (bits2 + num_v * 16) + 64 + (partial ? 32u : 0u);
            if (bytes_shifted) min_bits += n * shift * 2;
            do_escape = (min_bits >= escape_bits);
        }
        if (!do_escape) {
            /* post-check (:537-543) */
            min_bits = (uint32_t)(b->pos - start);
            if (min_bits >= escape_bits) do_escape = 2;
        }
        if (do_escape) {
            b->pos = start;
            if (tr) tr->escape = 2;
        }
    }
    if (do_escape)
        write_stereo_escape(e, b, pcm, stride, n);
    return ORC_OK;
}

/* codec/ALACEncoder.cu:812-963 EncodeMono */
static int32_t encode_mono(orc_encoder *e, orc_bits *b, const uint8_t *pcm, uint32_t stride,
                           uint32_t channel_index, uint32_t n, orc_trace *tr)
{
    const orc_prims *P = e->prims;
    int16_t (*cu)[MAX_COEFS] = e->coefs_u[channel_index];
    const uint32_t bytes_shifted = shift_bytes_for(e->bit_depth);
    const uint32_t shift = bytes_shifted * 8;
    const uint32_t chan_bits = e->bit_depth - shift;                          /* :857 */
    const uint32_t partial = (n == e->frame_size) ? 0u : 1u;
    const uint32_t clear_to = 32;   /* entries zeroed past the mixed count */
    const uint64_t start = b->pos;
    uint32_t min_bits = 1u << 31, best_u = MIN_UV, bits1 = 0;
    orc_ag_params ag;

    copy_mono(e, pcm, stride, n, shift, e->mix_u, e->shift_uv, clear_to);

    /* :881-905 */
    for (uint32_t nu = MIN_UV; nu <= MAX_UV; nu += 4) {
        for (int pass = 0; pass < 7; pass++)
            P->predict_enc(e->mix_u, e->pred_u, (int32_t)(n / 32), cu[nu - 1], (int32_t)nu, chan_bits, DENSHIFT_DEFAULT);
        P->predict_enc(e->mix_u, e->pred_u, (int32_t)(n / 8), cu[nu - 1], (int32_t)nu, chan_bits, DENSHIFT_DEFAULT);
        bits1 = trial_bits(e, e->pred_u, n / 8, chan_bits);
        uint32_t cost = 8 * bits1 + 16 * nu;
        if (cost < min_bits) { best_u = nu; min_bits = cost; }
    }
    /* :909-915 */
    min_bits += 32 + (partial ? 32u : 0u);
    if (bytes_shifted) min_bits += n * shift;
    const uint32_t escape_bits = (n * e->bit_depth) + (partial ? 32u : 0u) + 16u;
    int do_escape = (min_bits >= escape_bits);

    if (tr) {
        memset(tr, 0, sizeof(*tr));
        tr->tag = ORC_ID_SCE;
        tr->num_u = (int32_t)best_u;
        memcpy(tr->hdr_coefs_u, cu[best_u - 1], 8 * sizeof(int16_t));
        tr->escape = do_escape;
    }

    if (!do_escape) {
        /* :920-931 */
        orc_put(b, 0, 12);
        orc_put(b, (partial << 3) | (bytes_shifted << 1), 4);
        if (partial) orc_put(b, n, 32);
        orc_put(b, 0, 16);
        write_channel_header(b, best_u, cu[best_u - 1]);
        if (bytes_shifted)                                                    /* :934-938 */
            for (uint32_t i = 0; i < n; i++) orc_put(b, e->shift_uv[i], shift);
        /* :941-945 */
        P->predict_enc(e->mix_u, e->pred_u, (int32_t)n, cu[best_u - 1], (int32_t)best_u, chan_bits, DENSHIFT_DEFAULT);
        orc_ag_params_set(&ag, MB0, PB0, KB0);
        int32_t st = P->golomb_enc(&ag, e->pred_u, b, (int32_t)n, (int32_t)chan_bits, &bits1);
        if (st) return st;
        if (tr) tr->bits_u = bits1;
        /* :952-958 */
        min_bits = (uint32_t)(b->pos - start);
        if (min_bits >= escape_bits) {
            b->pos = start;
            do_escape = 2;
            if (tr) tr->escape = 2;
        }
    }
    if (do_escape)
        write_mono_escape(e, b, pcm, stride, n);
    return ORC_OK;
}

/* codec/ALACEncoder.cu:973-1057 Encode.  The >2-channel loop was deleted from the fork (F5);
   it is restated from sChannelMaps (:78-107) and the decoder's element loop
   (codec/ALACDecoder.cu:612-970): per-type instance tags count up from 0, stride = channels. */
int32_t orc_encode_packet(orc_encoder *e, const uint8_t *pcm, uint32_t num_samples,
                          uint8_t *out, uint32_t *out_bytes, orc_trace *trace)
{
    if (num_samples > e->frame_size) return ORC_PARAM_ERROR;
    const uint32_t bps = bytes_per_sample(e->bit_depth);
    orc_bits b;
    orc_bits_init(&b, out, (uint64_t)e->frame_size * e->channels * 5 + 64);   /* mMaxOutputBytes, :1489 */
    int32_t st = ORC_OK;
    uint32_t mono_tag = 0, stereo_tag = 0, lfe_tag = 0, ti = 0;

    if (e->channels == 1) {
        orc_put(&b, ORC_ID_SCE, 3);
        orc_put(&b, 0, 4);
        st = encode_mono(e, &b, pcm, 1, 0, num_samples, trace);
    } else if (e->channels == 2) {
        orc_put(&b, ORC_ID_CPE, 3);
        orc_put(&b, 0, 4);
        st = encode_pair(e, &b, pcm, 2, 0, num_samples, trace);
    } else {
        uint32_t ch = 0;
        while (ch < e->channels && st == ORC_OK) {
            uint32_t tag = (g_channel_maps[e->channels - 1] >> (ch * 3)) & 7u;
            orc_put(&b, tag, 3);
            orc_trace *tr = trace ? &trace[ti++] : NULL;
            switch (tag) {
            case ORC_ID_SCE:
                orc_put(&b, mono_tag++, 4);
                st = encode_mono(e, &b, pcm + (size_t)ch * bps, e->channels, ch, num_samples, tr);
                ch += 1;
                break;
            case ORC_ID_CPE:
                orc_put(&b, stereo_tag++, 4);
                st = encode_pair(e, &b, pcm + (size_t)ch * bps, e->channels, ch, num_samples, tr);
                ch += 2;
                break;
            case ORC_ID_LFE:
                orc_put(&b, lfe_tag++, 4);
                st = encode_mono(e, &b, pcm + (size_t)ch * bps, e->channels, ch, num_samples, tr);
                ch += 1;
                break;
            default:
                st = ORC_PARAM_ERROR;
            }
        }
    }
    if (st) return st;

    orc_put(&b, ORC_ID_END, 3);                                               /* :1036 */
    if (b.pos & 7u) orc_put(&b, 0, 8u - (unsigned)(b.pos & 7u));              /* :1039 */
    *out_bytes = (uint32_t)(b.pos >> 3);
    e->total_bytes += *out_bytes;                                             /* :1050-1051 */
    if (*out_bytes > e->max_frame_bytes) e->max_frame_bytes = *out_bytes;
    return ORC_OK;
}

int32_t orc_encode_stream(orc_encoder *e, const uint8_t *pcm, uint64_t num_sample_frames,
                          uint32_t frames_per_segment, uint8_t *out, uint64_t out_cap,
                          uint64_t *out_bytes, uint32_t *sizes, uint64_t *num_packets)
{
    const uint64_t bpf = (uint64_t)bytes_per_sample(e->bit_depth) * e->channels;
    uint64_t done = 0, written = 0, pkt = 0;
    uint8_t *tmp = (uint8_t *)malloc((size_t)e->frame_size * e->channels * 5 + 128);   /* worst pre-escape size, :1489 */
    if (!tmp) return ORC_MEM_ERROR;
    int32_t st = ORC_OK;
    while (done < num_sample_frames) {
        uint32_t n = (uint32_t)((num_sample_frames - done) < e->frame_size ? (num_sample_frames - done) : e->frame_size);
        if (frames_per_segment && (pkt % frames_per_segment) == 0) orc_encoder_reset(e);
        uint32_t nbytes = 0;
        st = orc_encode_packet(e, pcm + done * bpf, n, tmp, &nbytes, NULL);
        if (st) break;
        if (written + nbytes > out_cap) { st = ORC_PARAM_ERROR; break; }
        memcpy(out + written, tmp, nbytes);
        sizes[pkt++] = nbytes;
        written += nbytes;
        done += n;
    }
    free(tmp);
    *out_bytes = written;
    *num_packets = pkt;
    return st;
}

/* ------------------------------------------------------------------------- */
/* decoder                                                                    */
/* ------------------------------------------------------------------------- */
struct orc_decoder {
    orc_config cfg;
    const orc_prims *prims;
    int32_t *mix_u, *mix_v, *pred;
    uint16_t *shift_buf;
    uint8_t *padded;
    size_t padded_cap, padded_dirty;
};

static inline uint32_t get_be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

/* codec/ALACDecoder.cu:109-190 Init */
orc_decoder *orc_decoder_new(const uint8_t *cookie, uint32_t cookie_size, int flavour, int32_t *status)
{
    int32_t st_local;
    if (!status) status = &st_local;
    *status = ORC_PARAM_ERROR;
    const orc_prims *prims = pick_prims(flavour);
    if (!prims || !cookie) return NULL;
    const uint8_t *p = cookie;
    uint32_t left = cookie_size;
    if (left >= 12 && p[4] == 'f' && p[5] == 'r' && p[6] == 'm' && p[7] == 'a') { p += 12; left -= 12; }   /* :123-127 */
    if (left >= 12 && p[4] == 'a' && p[5] == 'l' && p[6] == 'a' && p[7] == 'c') { p += 12; left -= 12; }   /* :130-134 */
    if (left < 24) return NULL;                                                                         /* :137, :176-179 */

    orc_decoder *d = (orc_decoder *)calloc(1, sizeof(*d));
    if (!d) { *status = ORC_MEM_ERROR; return NULL; }
    d->prims = prims;
    d->cfg.frame_length = get_be32(p);
    d->cfg.compatible_version = p[4];
    d->cfg.bit_depth = p[5];
    d->cfg.pb = p[6];
    d->cfg.mb = p[7];
    d->cfg.kb = p[8];
    d->cfg.num_channels = p[9];
    d->cfg.max_run = (uint16_t)((p[10] << 8) | p[11]);
    d->cfg.max_frame_bytes = get_be32(p + 12);
    d->cfg.avg_bit_rate = get_be32(p + 16);
    d->cfg.sample_rate = get_be32(p + 20);
    if (d->cfg.compatible_version > 0) { free(d); return NULL; }                                        /* :153 */
    size_t n = (size_t)d->cfg.frame_length + 64;
    d->mix_u = (int32_t *)calloc(n, sizeof(int32_t));
    d->mix_v = (int32_t *)calloc(n, sizeof(int32_t));
    d->pred = (int32_t *)calloc(n, sizeof(int32_t));
    d->shift_buf = (uint16_t *)calloc(2 * n, sizeof(uint16_t));
    if (!d->mix_u || !d->mix_v || !d->pred || !d->shift_buf) { orc_decoder_free(d); *status = ORC_MEM_ERROR; return NULL; }
    *status = ORC_OK;
    return d;
}

void orc_decoder_free(orc_decoder *d)
{
    if (!d) return;
    free(d->mix_u); free(d->mix_v); free(d->pred); free(d->shift_buf); free(d->padded);
    free(d);
}

const orc_config *orc_decoder_config(const orc_decoder *d) { return &d->cfg; }

/* output stage: codec/ALACDecoder.cu:193-383 (unmixNN) and :385-495 (copyPredictorToNN).
   Follows Apple semantics where the fork differs: the shift merge is applied only when
   bytesShifted != 0 and every element uses its own channel offset. */
static void store_sample(uint8_t *pcm, uint32_t depth, uint64_t idx, int32_t val)
{
    uint8_t *p;
    switch (depth) {
    case 16:
        p = pcm + idx * 2;
        p[0] = (uint8_t)val; p[1] = (uint8_t)(val >> 8);
        break;
    case 20:
        val = (int32_t)((uint32_t)val << 4);
        /* fallthrough */
    case 24:
        p = pcm + idx * 3;
        p[0] = (uint8_t)val; p[1] = (uint8_t)(val >> 8); p[2] = (uint8_t)(val >> 16);
        break;
    default:
        p = pcm + idx * 4;
        p[0] = (uint8_t)val; p[1] = (uint8_t)(val >> 8); p[2] = (uint8_t)(val >> 16); p[3] = (uint8_t)(val >> 24);
    }
}

static int32_t read_channel_header(orc_bits *b, uint32_t *mode, uint32_t *den_shift, uint32_t *pb_factor,
                                   uint32_t *num, int16_t *coefs)
{
    uint32_t hb = orc_get(b, 8);                       /* codec/ALACDecoder.cu:660-669 */
    *mode = hb >> 4;
    *den_shift = hb & 0xfu;
    hb = orc_get(b, 8);
    *pb_factor = hb >> 5;
    *num = hb & 0x1fu;
    for (uint32_t i = 0; i < *num; i++) coefs[i] = (int16_t)orc_get(b, 16);
    return ORC_OK;
}

static int32_t decode_channel(orc_decoder *d, orc_bits *b, uint32_t n, uint32_t chan_bits, uint32_t mode,
                              uint32_t den_shift, uint32_t pb_factor, uint32_t num, int16_t *coefs, int32_t *dst)
{
    orc_ag_params ag;
    uint32_t bits = 0;
    /* codec/ALACDecoder.cu:682-694 */
    orc_ag_params_set(&ag, d->cfg.mb, ((uint32_t)d->cfg.pb * pb_factor) / 4, d->cfg.kb);
    int32_t st = d->prims->golomb_dec(&ag, b, d->pred, (int32_t)n, (int32_t)chan_bits, &bits);
    if (st) return st;
    if (mode == 0) {
        d->prims->predict_dec(d->pred, dst, (int32_t)n, coefs, (int32_t)num, chan_bits, den_shift);
    } else {
        d->prims->predict_dec(d->pred, d->pred, (int32_t)n, NULL, 31, chan_bits, 0);
        d->prims->predict_dec(d->pred, dst, (int32_t)n, coefs, (int32_t)num, chan_bits, den_shift);
    }
    return ORC_OK;
}

/* escape samples: codec/ALACDecoder.cu:697-727 */
static inline int32_t read_raw(orc_bits *b, uint32_t chan_bits)
{
    uint32_t sh = 32 - chan_bits;
    return (int32_t)(orc_get(b, chan_bits) << sh) >> sh;
}

/* codec/ALACDecoder.cu:571-1002 Decode */
int32_t orc_decode_packet(orc_decoder *d, const uint8_t *packet, uint32_t packet_bytes,
                          uint8_t *pcm_out, uint32_t *out_num_samples)
{
    const uint32_t depth = d->cfg.bit_depth;
    const uint32_t nch = d->cfg.num_channels;
    uint32_t n = d->cfg.frame_length;
    uint32_t channel_index = 0;
    int16_t coefs_u[32], coefs_v[32];
    int32_t st = ORC_OK;
    orc_bits b;

    if (nch == 0) return ORC_PARAM_ERROR;
    if (!(depth == 16 || depth == 20 || depth == 24 || depth == 32)) return ORC_PARAM_ERROR;
    /* Private copy in which every byte past the packet reads as zero -- the defined behaviour of this oracle (and of
       the CUDA bit readers) where the reference reads on into whatever follows its input buffer: escape samples and
       header fields of a truncated packet.  The tail is sized for the longest over-read (a whole raw frame); only the
       bytes an earlier, longer packet left behind need clearing. */
    {
        const size_t need = (size_t)packet_bytes + (size_t)d->cfg.frame_length * nch * 4u + 64u;
        if (d->padded_cap < need) {
            free(d->padded);
            d->padded_cap = need + 4096;
            d->padded = (uint8_t *)calloc(d->padded_cap, 1);
            if (!d->padded) { d->padded_cap = 0; return ORC_MEM_ERROR; }
            d->padded_dirty = 0;
        }
        memcpy(d->padded, packet, packet_bytes);
        if (d->padded_dirty > packet_bytes) memset(d->padded + packet_bytes, 0, d->padded_dirty - packet_bytes);
        d->padded_dirty = packet_bytes;
    }
    orc_bits_init(&b, d->padded, packet_bytes);
    *out_num_samples = n;

    while (st == ORC_OK) {
        if (!((b.pos >> 3) < packet_bytes)) { st = ORC_PARAM_ERROR; break; }          /* :615 */
        uint32_t tag = orc_get(&b, 3);
        if (tag == ORC_ID_SCE || tag == ORC_ID_LFE) {
            (void)orc_get(&b, 4);                                                    /* instance tag */
            if (orc_get(&b, 12) != 0) { st = ORC_PARAM_ERROR; break; }               /* :633 */
            uint32_t hb = orc_get(&b, 4);
            uint32_t partial = hb >> 3;
            uint32_t bytes_shifted = (hb >> 1) & 3u;
            if (bytes_shifted == 3) { st = ORC_PARAM_ERROR; break; }                 /* :641 */
            uint32_t shift = bytes_shifted * 8;
            uint32_t escape = hb & 1u;
            uint32_t chan_bits = depth - shift;
            if (partial) { n = orc_get(&b, 16) << 16; n |= orc_get(&b, 16); }         /* :650-654 */
            if (n > d->cfg.frame_length) { st = ORC_PARAM_ERROR; break; }            /* buffers are frame_length long */
            orc_bits shift_cursor = b;
            if (!escape) {
                (void)orc_get(&b, 8);                                                /* mixBits */
                (void)orc_get(&b, 8);                                                /* mixRes  */
                uint32_t mode, den, pbf, num;
                read_channel_header(&b, &mode, &den, &pbf, &num, coefs_u);
                if (bytes_shifted) { shift_cursor = b; b.pos += (uint64_t)shift * n; }    /* :675-679 */
                st = decode_channel(d, &b, n, chan_bits, mode, den, pbf, num, coefs_u, d->mix_u);
                if (st) break;
            } else {
                for (uint32_t i = 0; i < n; i++) d->mix_u[i] = read_raw(&b, chan_bits);
                bytes_shifted = 0;
            }
            if (bytes_shifted)                                                      /* :730-737 */
                for (uint32_t i = 0; i < n; i++) d->shift_buf[i] = (uint16_t)orc_get(&shift_cursor, shift);
            if (channel_index < nch) {
                for (uint32_t i = 0; i < n; i++) {
                    int32_t v = d->mix_u[i];
                    if (bytes_shifted) v = (int32_t)(((uint32_t)v << shift) | d->shift_buf[i]);
                    store_sample(pcm_out, depth, (uint64_t)i * nch + channel_index, v);
                }
            }
            channel_index += 1;
            *out_num_samples = n;
        } else if (tag == ORC_ID_CPE) {
            if (channel_index + 2 > nch) break;                                      /* :759-760 */
            (void)orc_get(&b, 4);
            if (orc_get(&b, 12) != 0) { st = ORC_PARAM_ERROR; break; }
            uint32_t hb = orc_get(&b, 4);
            uint32_t partial = hb >> 3;
            uint32_t bytes_shifted = (hb >> 1) & 3u;
            if (bytes_shifted == 3) { st = ORC_PARAM_ERROR; break; }
            uint32_t shift = bytes_shifted * 8;
            uint32_t escape = hb & 1u;
            uint32_t chan_bits = depth - shift + 1;
            if (partial) { n = orc_get(&b, 16) << 16; n |= orc_get(&b, 16); }
            if (n > d->cfg.frame_length) { st = ORC_PARAM_ERROR; break; }
            uint32_t mix_bits = 0;
            int32_t mix_res = 0;
            orc_bits shift_cursor = b;
            if (!escape) {
                mix_bits = orc_get(&b, 8);
                mix_res = (int8_t)orc_get(&b, 8);
                uint32_t mode_u, den_u, pbf_u, num_u, mode_v, den_v, pbf_v, num_v;
                read_channel_header(&b, &mode_u, &den_u, &pbf_u, &num_u, coefs_u);
                read_channel_header(&b, &mode_v, &den_v, &pbf_v, &num_v, coefs_v);
                if (bytes_shifted) { shift_cursor = b; b.pos += (uint64_t)shift * 2 * n; }    /* :818-822 */
                st = decode_channel(d, &b, n, chan_bits, mode_u, den_u, pbf_u, num_u, coefs_u, d->mix_u);
                if (st) break;
                st = decode_channel(d, &b, n, chan_bits, mode_v, den_v, pbf_v, num_v, coefs_v, d->mix_v);
                if (st) break;
            } else {
                chan_bits = depth;                                                  /* :858 */
                for (uint32_t i = 0; i < n; i++) {
                    d->mix_u[i] = read_raw(&b, chan_bits);
                    d->mix_v[i] = read_raw(&b, chan_bits);
                }
                bytes_shifted = 0;
            }
            if (bytes_shifted)                                                      /* :899-909 */
                for (uint32_t i = 0; i < 2 * n; i++) d->shift_buf[i] = (uint16_t)orc_get(&shift_cursor, shift);
            for (uint32_t i = 0; i < n; i++) {                                       /* :193-383 */
                int32_t l, r;
                if (mix_res != 0) {
                    l = d->mix_u[i] + d->mix_v[i] - ((mix_res * d->mix_v[i]) >> mix_bits);
                    r = l - d->mix_v[i];
                } else {
                    l = d->mix_u[i];
                    r = d->mix_v[i];
                }
                if (bytes_shifted) {
                    l = (int32_t)(((uint32_t)l << shift) | d->shift_buf[2 * i]);
                    r = (int32_t)(((uint32_t)r << shift) | d->shift_buf[2 * i + 1]);
                }
                store_sample(pcm_out, depth, (uint64_t)i * nch + channel_index, l);
                store_sample(pcm_out, depth, (uint64_t)i * nch + channel_index + 1, r);
            }
            channel_index += 2;
            *out_num_samples = n;
        } else if (tag == ORC_ID_CCE || tag == ORC_ID_PCE) {
            st = ORC_PARAM_ERROR;                                                   /* :932-939 */
        } else if (tag == ORC_ID_DSE) {
            /* codec/ALACDecoder.cu:1033-1059 */
            (void)orc_get(&b, 4);
            uint32_t align = orc_get(&b, 1);
            uint32_t count = orc_get(&b, 8);
            if (count == 255) count += orc_get(&b, 8);
            if (align && (b.pos & 7u)) b.pos += 8u - (b.pos & 7u);
            b.pos += (uint64_t)count * 8;
            if ((b.pos >> 3) > packet_bytes) st = ORC_PARAM_ERROR;
        } else if (tag == ORC_ID_FIL) {
            /* codec/ALACDecoder.cu:1012-1027 */
            int32_t count = (int32_t)orc_get(&b, 4);
            if (count == 15) count += (int32_t)orc_get(&b, 8) - 1;
            b.pos += (uint64_t)count * 8;
            if ((b.pos >> 3) > packet_bytes) st = ORC_PARAM_ERROR;
        } else { /* ORC_ID_END :955-961 */
            if (b.pos & 7u) b.pos += 8u - (b.pos & 7u);
            break;
        }
        if (channel_index >= nch) break;                                             /* :966-967 */
    }

    /* :972-998 channels that never arrived are zero-filled (Apple; commented out in the fork) */
    if (st == ORC_OK)
        for (; channel_index < nch; channel_index++)
            for (uint32_t i = 0; i < *out_num_samples; i++)
                store_sample(pcm_out, depth, (uint64_t)i * nch + channel_index, 0);
    return st;
}

int32_t orc_decode_stream(orc_decoder *d, const uint8_t *packets, const uint32_t *sizes,
                          uint64_t num_packets, uint8_t *pcm_out, uint64_t pcm_cap,
                          uint64_t *out_sample_frames, int32_t *statuses)
{
    const uint64_t bpf = (uint64_t)bytes_per_sample(d->cfg.bit_depth) * d->cfg.num_channels;
    uint64_t in_off = 0, frames = 0;
    int32_t first_err = ORC_OK;
    uint8_t *tmp = (uint8_t *)malloc((size_t)d->cfg.frame_length * bpf + 64);
    if (!tmp) return ORC_MEM_ERROR;
    for (uint64_t i = 0; i < num_packets; i++) {
        uint32_t n = 0;
        int32_t st = orc_decode_packet(d, packets + in_off, sizes[i], tmp, &n);
        if (statuses) statuses[i] = st;
        if (st && !first_err) first_err = st;
        if (st == ORC_OK) {
            if ((frames + n) * bpf > pcm_cap) { first_err = ORC_PARAM_ERROR; break; }
            memcpy(pcm_out + frames * bpf, tmp, (size_t)n * bpf);
            frames += n;
        }
        in_off += sizes[i];
    }
    free(tmp);
    *out_sample_frames = frames;
    return first_err;
}
