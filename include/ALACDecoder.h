// ALACDecoder.h -- drop-in ALACDecoder over the B200 engine.
// Public surface of the reference class (codec/ALACDecoder.h:38-72).  libalac's signatures
// (Init(cookie, size), Decode(bits, sampleBuffer, numSamples, numChannels, outNumSamples)) and the
// fork's (`int X` slot arguments + fillWriteBuffer) are both provided.
#ifndef ALACDECODER_H
#define ALACDECODER_H
#include <stdint.h>
#include <vector>
#include "ALACAudioTypes.h"
#include "alac_b200.h"

struct BitBuffer;

class ALACDecoder {
public:
    ALACDecoder();
    ~ALACDecoder();

    int32_t Init(void *inMagicCookie, uint32_t inMagicCookieSize);
    int32_t Init(void *inMagicCookie, uint32_t inMagicCookieSize, int X);
    // libalac: decode one packet, interleaved PCM written to sampleBuffer (host)
    int32_t Decode(struct BitBuffer *bits, uint8_t *sampleBuffer, uint32_t numSamples, uint32_t numChannels,
                   uint32_t *outNumSamples);
    // fork: decode packet X into the object's staging area; fillWriteBuffer() then lays all staged
    // packets out at X * theOutputPacketBytes in a caller-provided DEVICE buffer (main.cu:717-747)
    int32_t Decode(struct BitBuffer *bits, uint32_t numSamples, uint32_t numChannels, uint32_t *outNumSamples,
                   uint32_t outBytesPerPacket, int X);
    void fillWriteBuffer(void *sampleBuffer, uint32_t numChannels, int32_t theOutputPacketBytes, int X);

    // Batched extension: decode packets laid back to back (host buffers) in one call.
    int32_t DecodeBatch(const unsigned char *packets, const uint32_t *packetSizes, uint64_t numPackets,
                        unsigned char *pcmOut, uint64_t pcmCap, uint64_t *outSampleFrames, int32_t *packetStatus);

public:
    ALACSpecificConfig mConfig;      // host-endian copy of the cookie, as in the reference

private:
    alac_b200_engine *mEngine;
    uint8_t mCookie[ALAC_B200_COOKIE_MAX];
    uint32_t mCookieSize;
    std::vector<std::vector<uint8_t> > mStaged;     // fork path: decoded packet X
};
#endif
