/*
 * alac_oracle.h -- CPU oracle for the ALAC encode/decode hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker or as the timed CPU arm.
 *
 * The oracle is a plain-C restatement of the reference's algorithm (file:line
 * citations are relative to /root/reference).  Its four hot primitives
 * (predictor encode/decode, adaptive-Golomb encode/decode) are reached through a
 * function table so the same frame drivers can run on
 *   - the restated primitives in alac_oracle.c               ("port"), or
 *   - the reference's own unmodified dp_enc.c / dp_dec.c / ag_enc.c / ag_dec.c,
 *     compiled from /root/reference into oracle/_ref/ (ref_adapter.c)  ("reference").
 * Parity pinning: tests/test_oracle.py checks the restated primitives against
 * SURVEY.md Appendix D known-answer vectors (generated from the reference's own
 * objects), against the reference objects in oracle/_ref on random inputs, and
 * against tests/golden/ packet fixtures produced through the reference primitives.
 */
#ifndef ALAC_ORACLE_H
#define ALAC_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes: codec/ALACAudioTypes.h:54-60, codec/ALACBitUtilities.h:51-54 */
#define ORC_OK             0
#define ORC_UNIMPLEMENTED  (-4)
#define ORC_PARAM_ERROR    (-50)
#define ORC_MEM_ERROR      (-108)

/* element tags: codec/ALACBitUtilities.h:57-68 */
enum { ORC_ID_SCE = 0, ORC_ID_CPE = 1, ORC_ID_CCE = 2, ORC_ID_LFE = 3,
       ORC_ID_DSE = 4, ORC_ID_PCE = 5, ORC_ID_FIL = 6, ORC_ID_END = 7 };

/* ---- MSB-first bit cursor over a caller-owned byte buffer ---------------- */
typedef struct {
    uint8_t *buf;       /* start of buffer                                   */
    uint64_t pos;       /* current bit position from buf                     */
    uint64_t cap;       /* capacity in bits                                  */
} orc_bits;

void     orc_bits_init(orc_bits *b, uint8_t *buf, uint64_t cap_bytes);
void     orc_put(orc_bits *b, uint32_t value, unsigned nbits);        /* low nbits of value, nbits 0..32 */
uint32_t orc_get(orc_bits *b, unsigned nbits);                        /* nbits 0..32 */

/* ---- adaptive Golomb parameters: codec/aglib.h:57-66 ---------------------- */
typedef struct {
    uint32_t mb0, pb, kb, wb;
} orc_ag_params;

void orc_ag_params_set(orc_ag_params *p, uint32_t mb0, uint32_t pb, uint32_t kb);

/* ---- primitive table ------------------------------------------------------ */
typedef struct {
    const char *name;
    /* codec/dp_enc.c:77 pc_block */
    void (*predict_enc)(const int32_t *in, int32_t *res, int32_t num, int16_t *coefs,
                        int32_t numactive, uint32_t chanbits, uint32_t denshift);
    /* codec/dp_dec.c:55 unpc_block */
    void (*predict_dec)(const int32_t *res, int32_t *out, int32_t num, int16_t *coefs,
                        int32_t numactive, uint32_t chanbits, uint32_t denshift);
    /* codec/ag_enc.c:249 dyn_comp; appends at b->pos, returns status, *out_bits = bits written */
    int32_t (*golomb_enc)(const orc_ag_params *p, const int32_t *res, orc_bits *b,
                          int32_t num, int32_t bit_size, uint32_t *out_bits);
    /* codec/ag_dec.c:272 dyn_decomp; reads at b->pos (b->cap = packet bits) */
    int32_t (*golomb_dec)(const orc_ag_params *p, orc_bits *b, int32_t *res,
                          int32_t num, int32_t max_size, uint32_t *out_bits);
} orc_prims;

const orc_prims *orc_prims_port(void);            /* restated primitives (always available) */
/* installed by the _ref flavour (ref_adapter.c); NULL in the plain build */
const orc_prims *orc_prims_reference(void);

/* restated primitives, exported individually for KAT tests */
void    orc_init_coefs(int16_t *coefs, uint32_t denshift, int32_t n);
void    orc_pc_block(const int32_t *in, int32_t *res, int32_t num, int16_t *coefs,
                     int32_t numactive, uint32_t chanbits, uint32_t denshift);
void    orc_unpc_block(const int32_t *res, int32_t *out, int32_t num, int16_t *coefs,
                       int32_t numactive, uint32_t chanbits, uint32_t denshift);
int32_t orc_dyn_comp(const orc_ag_params *p, const int32_t *res, orc_bits *b,
                     int32_t num, int32_t bit_size, uint32_t *out_bits);
int32_t orc_dyn_decomp(const orc_ag_params *p, orc_bits *b, int32_t *res,
                       int32_t num, int32_t max_size, uint32_t *out_bits);

/* ---- per-element trace (debug aid for GPU mismatches) ---------------------- */
typedef struct {
    int32_t  tag;            /* ORC_ID_SCE / ORC_ID_CPE                              */
    int32_t  escape;         /* 0 compressed, 1 escape by estimate, 2 by post-check  */
    int32_t  mix_res;
    int32_t  num_u, num_v;
    uint32_t bits_u, bits_v; /* final-pass Golomb bits                               */
    int16_t  hdr_coefs_u[8], hdr_coefs_v[8];
} orc_trace;

/* ---- encoder --------------------------------------------------------------- */
typedef struct orc_encoder orc_encoder;

/* flavour: 0 = port primitives, 1 = reference primitives (fails -> NULL if not linked) */
orc_encoder *orc_encoder_new(uint32_t channels, uint32_t bit_depth, uint32_t sample_rate,
                             uint32_t frame_size, int fast_mode, int flavour);
void     orc_encoder_free(orc_encoder *e);
/* re-run init_coefs on every row: the state a fresh ALACEncoder starts from
   (codec/ALACEncoder.cu:1524-1531).  Used at segment boundaries. */
void     orc_encoder_reset(orc_encoder *e);
/* codec/ALACEncoder.cu:1109-1140; returns cookie size or 0 if cap too small */
uint32_t orc_encoder_cookie(const orc_encoder *e, uint8_t *out, uint32_t cap);
/* codec/ALACEncoder.cu:973-1057 (+ multichannel loop via sChannelMaps :97-107).
   pcm: interleaved little-endian packed PCM, num_samples sample-frames (<= frame_size).
   out must hold frame_size*channels*5 + 64 bytes (the pre-escape worst case, mMaxOutputBytes).
   trace may be NULL (else >= 8 entries). */
int32_t  orc_encode_packet(orc_encoder *e, const uint8_t *pcm, uint32_t num_samples,
                           uint8_t *out, uint32_t *out_bytes, orc_trace *trace);
/* whole stream: frames of frame_size, tail partial; encoder reset every frames_per_segment
   frames (0 = never).  sizes[] gets one entry per packet.  Returns status. */
int32_t  orc_encode_stream(orc_encoder *e, const uint8_t *pcm, uint64_t num_sample_frames,
                           uint32_t frames_per_segment, uint8_t *out, uint64_t out_cap,
                           uint64_t *out_bytes, uint32_t *sizes, uint64_t *num_packets);
/* encoder coefficient state export/import (rows 3 and 7 only are live) */
void     orc_encoder_get_coefs(const orc_encoder *e, uint32_t channel, int is_v, uint32_t row, int16_t *out16);

/* ---- decoder --------------------------------------------------------------- */
typedef struct {
    uint32_t frame_length;
    uint8_t  compatible_version, bit_depth, pb, mb, kb, num_channels;
    uint16_t max_run;
    uint32_t max_frame_bytes, avg_bit_rate, sample_rate;
} orc_config;

typedef struct orc_decoder orc_decoder;

orc_decoder *orc_decoder_new(const uint8_t *cookie, uint32_t cookie_size, int flavour, int32_t *status);
void     orc_decoder_free(orc_decoder *d);
const orc_config *orc_decoder_config(const orc_decoder *d);
/* codec/ALACDecoder.cu:571-1002 with the output stage of :193-495 applied per element.
   pcm_out: interleaved LE packed PCM, needs frame_length*channels*bytes. */
int32_t  orc_decode_packet(orc_decoder *d, const uint8_t *packet, uint32_t packet_bytes,
                           uint8_t *pcm_out, uint32_t *out_num_samples);
int32_t  orc_decode_stream(orc_decoder *d, const uint8_t *packets, const uint32_t *sizes,
                           uint64_t num_packets, uint8_t *pcm_out, uint64_t pcm_cap,
                           uint64_t *out_sample_frames, int32_t *statuses);

uint32_t orc_fnv1a(const void *data, size_t n);

#ifdef __cplusplus
}
#endif
#endif
