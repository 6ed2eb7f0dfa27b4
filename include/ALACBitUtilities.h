// ALACBitUtilities.h -- the BitBuffer handle ALACDecoder::Decode takes (reference:
// codec/ALACBitUtilities.h:71-78 struct, :57-68 element tags, :51-54 ALAC_noErr).
// Only what callers of the class API touch is provided; the codec itself never runs on the host.
#ifndef ALACBITUTILITIES_H
#define ALACBITUTILITIES_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ALAC_noErr = 0 };
typedef enum { ID_SCE = 0, ID_CPE = 1, ID_CCE = 2, ID_LFE = 3, ID_DSE = 4, ID_PCE = 5, ID_FIL = 6, ID_END = 7 } ELEMENT_TYPE;

typedef struct BitBuffer {
    uint8_t *cur;
    uint8_t *end;
    uint32_t bitIndex;
    uint32_t byteSize;
} BitBuffer;

static inline void BitBufferInit(BitBuffer *bits, uint8_t *buffer, uint32_t byteSize)
{
    bits->cur = buffer;
    bits->end = buffer + byteSize;
    bits->bitIndex = 0;
    bits->byteSize = byteSize;
}
static inline uint32_t BitBufferGetPosition(const BitBuffer *bits)
{
    return (uint32_t)(bits->cur - (bits->end - bits->byteSize)) * 8u + bits->bitIndex;
}
static inline void BitBufferReset(BitBuffer *bits)
{
    bits->cur = bits->end - bits->byteSize;
    bits->bitIndex = 0;
}

#ifdef __cplusplus
}
#endif
#endif
