/*
 * ref_adapter.c -- binds the oracle's primitive table to the REFERENCE's own,
 * unmodified pc_block / unpc_block / dyn_comp / dyn_decomp.
 *
 * TEST INFRASTRUCTURE ONLY.  Compiled only into oracle/_ref/liboracle_ref.so by
 * oracle/Makefile, which compiles the reference's C sources where they lie under
 * /root/reference/codec (no reference source is copied into this repo):
 *   dp_enc.c dp_dec.c ag_enc.c ag_dec.c ALACBitUtilities.c EndianPortable.c
 * The reference's class drivers (ALACEncoder.cu / ALACDecoder.cu) need a GPU at
 * run time and are therefore restated in alac_oracle.c; this adapter lets those
 * restated drivers run on the reference's real arithmetic.
 */
#include "alac_oracle.h"

#include "aglib.h"               /* /root/reference/codec/aglib.h  */
#include "dplib.h"               /* /root/reference/codec/dplib.h  */
#include "ALACBitUtilities.h"    /* /root/reference/codec/ALACBitUtilities.h */

static void ref_predict_enc(const int32_t *in, int32_t *res, int32_t num, int16_t *coefs,
                            int32_t numactive, uint32_t chanbits, uint32_t denshift)
{
    pc_block((int32_t *)in, res, num, coefs, numactive, chanbits, denshift);
}

static void ref_predict_dec(const int32_t *res, int32_t *out, int32_t num, int16_t *coefs,
                            int32_t numactive, uint32_t chanbits, uint32_t denshift)
{
    unpc_block((int32_t *)res, out, num, coefs, numactive, chanbits, denshift);
}

static void cursor_to_bitbuffer(const orc_bits *b, BitBuffer *bb)
{
    bb->byteSize = (uint32_t)(b->cap >> 3);
    bb->end = b->buf + bb->byteSize;
    bb->cur = b->buf + (b->pos >> 3);
    bb->bitIndex = (uint32_t)(b->pos & 7u);
}

static int32_t ref_golomb_enc(const orc_ag_params *p, const int32_t *res, orc_bits *b,
                              int32_t num, int32_t bit_size, uint32_t *out_bits)
{
    AGParamRec ag;
    BitBuffer bb;
    set_ag_params(&ag, p->mb0, p->pb, p->kb, (uint32_t)num, (uint32_t)num, MAX_RUN_DEFAULT);
    cursor_to_bitbuffer(b, &bb);
    int32_t st = dyn_comp(&ag, (int32_t *)res, &bb, num, bit_size, out_bits);
    b->pos += *out_bits;
    return st;
}

static int32_t ref_golomb_dec(const orc_ag_params *p, orc_bits *b, int32_t *res,
                              int32_t num, int32_t max_size, uint32_t *out_bits)
{
    AGParamRec ag;
    BitBuffer bb;
    set_ag_params(&ag, p->mb0, p->pb, p->kb, (uint32_t)num, (uint32_t)num, MAX_RUN_DEFAULT);
    cursor_to_bitbuffer(b, &bb);
    int32_t st = dyn_decomp(&ag, &bb, res, num, max_size, out_bits);
    b->pos += *out_bits;
    return st;
}

static const orc_prims g_ref_prims = {
    "reference", ref_predict_enc, ref_predict_dec, ref_golomb_enc, ref_golomb_dec
};

const orc_prims *orc_prims_reference(void) { return &g_ref_prims; }

/* raw entry points for primitive-vs-primitive tests */
void orc_ref_init_coefs(int16_t *coefs, uint32_t denshift, int32_t n) { init_coefs(coefs, denshift, n); }
