// ALACEncoder.h -- drop-in ALACEncoder over the B200 engine.
// Public surface of the reference class (codec/ALACEncoder.h:34-102): same method names, argument
// meaning and int32 status codes.  Both signature families are provided: libalac's
// (InitializeEncoder(fmt), Encode(in, out, read, write, ioNumBytes)) and the fork's with the extra
// `int X` / `int index` arguments and InitializeSampling (accepted and ignored).
// Each Encode() call runs one frame through the sm_100a kernels with the predictor state carried
// between calls exactly like mCoefsU/mCoefsV; for throughput use EncodeBatch / alac_b200_encode.
#ifndef ALACENCODER_H
#define ALACENCODER_H
#include <stdint.h>
#include "ALACAudioTypes.h"
#include "alac_b200.h"

class ALACEncoder {
public:
    ALACEncoder();
    virtual ~ALACEncoder();

    virtual int32_t Encode(AudioFormatDescription theInputFormat, AudioFormatDescription theOutputFormat,
                           unsigned char *theReadBuffer, unsigned char *theWriteBuffer, int32_t *ioNumBytes);
    virtual int32_t Encode(AudioFormatDescription theInputFormat, AudioFormatDescription theOutputFormat,
                           unsigned char *theReadBuffer, unsigned char *theWriteBuffer, int32_t *ioNumBytes, int index);
    virtual int32_t Finish();

    void SetFastMode(bool fast) { mFastMode = fast; }
    void SetFrameSize(uint32_t frameSize) { mFrameSize = frameSize; }     // before InitializeEncoder()

    void GetConfig(ALACSpecificConfig &config);
    uint32_t GetMagicCookieSize(uint32_t inNumChannels);
    void GetMagicCookie(void *config, uint32_t *ioSize);

    virtual int32_t InitializeEncoder(AudioFormatDescription theOutputFormat);
    virtual int32_t InitializeEncoder(AudioFormatDescription theOutputFormat, int X);
    void InitializeSampling(void *d_ip, AudioFormatDescription theInputFormat, int X, int32_t *outBytes);

    // Batched extension: encode a whole host PCM buffer in one call.  framesPerSegment is the
    // encoder-reset schedule (0 = continue this object's state serially, like repeated Encode()).
    int32_t EncodeBatch(const unsigned char *pcm, uint64_t numSampleFrames, uint32_t framesPerSegment,
                        unsigned char *packetsOut, uint64_t packetsCap, uint32_t *packetSizes, uint64_t sizesCap,
                        uint64_t *outNumPackets, uint64_t *outBytes);

protected:
    int16_t mBitDepth;
    bool mFastMode;
    uint32_t mTotalBytesGenerated, mAvgBitRate, mMaxFrameBytes;
    uint32_t mFrameSize, mMaxOutputBytes, mNumChannels, mOutputSampleRate;

private:
    alac_b200_engine *mEngine;
    int16_t mCoefState[ALAC_B200_STATE_INT16S];     // live rows of mCoefsU/mCoefsV
    void ResetState();
    alac_b200_enc_config MakeConfig(uint32_t framesPerSegment) const;
};
#endif
