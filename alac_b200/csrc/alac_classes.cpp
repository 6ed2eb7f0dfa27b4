// alac_classes.cpp -- ALACEncoder / ALACDecoder (include/ALACEncoder.h, include/ALACDecoder.h):
// the reference's class API (codec/ALACEncoder.h:34-102, codec/ALACDecoder.h:38-72) implemented on
// the C ABI of libalac_b200.  All codec arithmetic runs in the CUDA kernels; nothing here encodes
// or decodes on the host.
#include "../../include/ALACBitUtilities.h"
#include "../../include/ALACDecoder.h"
#include "../../include/ALACEncoder.h"

#include <cstdio>
#include <cstring>

// ---------------------------------------------------------------------------------------------
// ALACEncoder
// ---------------------------------------------------------------------------------------------
ALACEncoder::ALACEncoder()
    : mBitDepth(0), mFastMode(false), mTotalBytesGenerated(0), mAvgBitRate(0), mMaxFrameBytes(0),
      mFrameSize(kALACDefaultFrameSize), mMaxOutputBytes(0), mNumChannels(0), mOutputSampleRate(0), mEngine(nullptr)
{
    ResetState();
}

ALACEncoder::~ALACEncoder()
{
    if (mEngine) alac_b200_engine_destroy(mEngine);
}

void ALACEncoder::ResetState()
{
    // init_coefs on every live row (codec/dp_enc.c:49-60, codec/ALACEncoder.cu:1524-1531)
    for (int row = 0; row < ALAC_B200_STATE_INT16S / 8; row++) {
        int16_t *c = mCoefState + row * 8;
        memset(c, 0, 8 * sizeof(int16_t));
        c[0] = (38 * 512) >> 4;
        c[1] = (-29 * 512) >> 4;
        c[2] = (-2 * 512) >> 4;
    }
}

alac_b200_enc_config ALACEncoder::MakeConfig(uint32_t framesPerSegment) const
{
    alac_b200_enc_config c;
    c.sample_rate = mOutputSampleRate;
    c.channels = mNumChannels;
    c.bit_depth = (uint32_t)mBitDepth;
    c.frame_size = mFrameSize;
    c.fast_mode = mFastMode ? 1u : 0u;
    c.frames_per_segment = framesPerSegment;
    return c;
}

int32_t ALACEncoder::InitializeEncoder(AudioFormatDescription theOutputFormat)
{
    // codec/ALACEncoder.cu:1457-1535
    mOutputSampleRate = (uint32_t)theOutputFormat.mSampleRate;
    mNumChannels = theOutputFormat.mChannelsPerFrame;
    switch (theOutputFormat.mFormatFlags) {
    case 1: mBitDepth = 16; break;
    case 2: mBitDepth = 20; break;
    case 3: mBitDepth = 24; break;
    case 4: mBitDepth = 32; break;
    default: break;
    }
    if (mNumChannels < 1 || mNumChannels > kALACMaxChannels || mBitDepth == 0 || mFrameSize == 0) return kALAC_ParamError;
    mMaxOutputBytes = mFrameSize * mNumChannels * ((10 + 32) / 8) + 1;
    ResetState();
    if (!mEngine) {
        int32_t st = alac_b200_engine_create(-1, &mEngine);
        if (st != ALAC_B200_OK) return st == ALAC_B200_MEM_ERROR ? kALAC_MemFullError : st;
    }
    return ALAC_noErr;
}

int32_t ALACEncoder::InitializeEncoder(AudioFormatDescription theOutputFormat, int /*X*/)
{
    return InitializeEncoder(theOutputFormat);
}

void ALACEncoder::InitializeSampling(void *, AudioFormatDescription, int, int32_t *)
{
    // the fork pre-mixes the whole file here (codec/ALACEncoder.cu:1385-1451); mixing is fused into
    // the search kernel, so there is nothing to prepare
}

int32_t ALACEncoder::Encode(AudioFormatDescription theInputFormat, AudioFormatDescription /*theOutputFormat*/,
                            unsigned char *theReadBuffer, unsigned char *theWriteBuffer, int32_t *ioNumBytes)
{
    // codec/ALACEncoder.cu:973-1057: one packet per call, state carried in mCoefState
    if (!mEngine || !ioNumBytes || !theReadBuffer || !theWriteBuffer) return kALAC_ParamError;
    if (theInputFormat.mBytesPerPacket == 0) return kALAC_ParamError;
    const uint32_t numFrames = (uint32_t)*ioNumBytes / theInputFormat.mBytesPerPacket;
    if (numFrames > mFrameSize) return kALAC_ParamError;
    const alac_b200_enc_config cfg = MakeConfig(0);
    // the caller's buffer must hold the escape-sized packet: input bytes + kALACMaxEscapeHeaderBytes
    const uint64_t cap = alac_b200_encode_bound(&cfg, numFrames, 1);
    uint32_t size = 0;
    uint64_t npk = 0, nbytes = 0;
    if (numFrames == 0) {
        // an empty call still yields ID_END padded to a byte (codec/ALACEncoder.cu:1036-1039)
        theWriteBuffer[0] = 0xE0;
        *ioNumBytes = 1;
        return ALAC_noErr;
    }
    int32_t st = alac_b200_encode(mEngine, &cfg, theReadBuffer, numFrames, ALAC_B200_MEM_HOST, nullptr, 1,
                                  theWriteBuffer, cap, &size, 1, ALAC_B200_MEM_HOST, mCoefState, &npk, &nbytes, nullptr);
    if (st != ALAC_B200_OK) return st;
    *ioNumBytes = (int32_t)nbytes;
    mTotalBytesGenerated += (uint32_t)nbytes;                               // :1050-1051
    if ((uint32_t)nbytes > mMaxFrameBytes) mMaxFrameBytes = (uint32_t)nbytes;
    return ALAC_noErr;
}

int32_t ALACEncoder::Encode(AudioFormatDescription theInputFormat, AudioFormatDescription theOutputFormat,
                            unsigned char *theReadBuffer, unsigned char *theWriteBuffer, int32_t *ioNumBytes, int /*index*/)
{
    return Encode(theInputFormat, theOutputFormat, theReadBuffer, theWriteBuffer, ioNumBytes);
}

int32_t ALACEncoder::EncodeBatch(const unsigned char *pcm, uint64_t numSampleFrames, uint32_t framesPerSegment,
                                 unsigned char *packetsOut, uint64_t packetsCap, uint32_t *packetSizes, uint64_t sizesCap,
                                 uint64_t *outNumPackets, uint64_t *outBytes)
{
    if (!mEngine) return kALAC_ParamError;
    const alac_b200_enc_config cfg = MakeConfig(framesPerSegment);
    alac_b200_stats stats;
    int32_t st = alac_b200_encode(mEngine, &cfg, pcm, numSampleFrames, ALAC_B200_MEM_HOST, nullptr, 1, packetsOut,
                                  packetsCap, packetSizes, sizesCap, ALAC_B200_MEM_HOST,
                                  framesPerSegment == 0 ? mCoefState : nullptr, outNumPackets, outBytes, &stats);
    if (st == ALAC_B200_OK) {
        mTotalBytesGenerated += (uint32_t)stats.payload_bytes;
        if (stats.max_packet_bytes > mMaxFrameBytes) mMaxFrameBytes = stats.max_packet_bytes;
    }
    return st;
}

int32_t ALACEncoder::Finish() { return ALAC_noErr; }                       // codec/ALACEncoder.cu:1064-1074

static inline uint32_t swap32(uint32_t v) { return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24); }
static inline uint16_t swap16(uint16_t v) { return (uint16_t)((v >> 8) | (v << 8)); }

void ALACEncoder::GetConfig(ALACSpecificConfig &config)
{
    // codec/ALACEncoder.cu:1082-1095: multi-byte fields are stored big-endian (little-endian host)
    config.frameLength = swap32(mFrameSize);
    config.compatibleVersion = (uint8_t)kALACCompatibleVersion;
    config.bitDepth = (uint8_t)mBitDepth;
    config.pb = 40;
    config.kb = 14;
    config.mb = 10;
    config.numChannels = (uint8_t)mNumChannels;
    config.maxRun = swap16(255);
    config.maxFrameBytes = swap32(mMaxFrameBytes);
    config.avgBitRate = swap32(mAvgBitRate);
    config.sampleRate = swap32(mOutputSampleRate);
}

uint32_t ALACEncoder::GetMagicCookieSize(uint32_t inNumChannels)
{
    return inNumChannels > 2 ? (uint32_t)(sizeof(ALACSpecificConfig) + kChannelAtomSize + sizeof(ALACAudioChannelLayout))
                             : (uint32_t)sizeof(ALACSpecificConfig);
}

void ALACEncoder::GetMagicCookie(void *outCookie, uint32_t *ioSize)
{
    // codec/ALACEncoder.cu:1109-1140
    const alac_b200_enc_config cfg = MakeConfig(0);
    *ioSize = alac_b200_magic_cookie(&cfg, mMaxFrameBytes, mAvgBitRate, outCookie, *ioSize);
}

// ---------------------------------------------------------------------------------------------
// ALACDecoder
// ---------------------------------------------------------------------------------------------
ALACDecoder::ALACDecoder() : mEngine(nullptr), mCookieSize(0) { memset(&mConfig, 0, sizeof(mConfig)); }

ALACDecoder::~ALACDecoder()
{
    if (mEngine) alac_b200_engine_destroy(mEngine);
}

int32_t ALACDecoder::Init(void *inMagicCookie, uint32_t inMagicCookieSize)
{
    // codec/ALACDecoder.cu:109-190
    uint32_t f[11];
    int32_t st = alac_b200_parse_cookie(inMagicCookie, inMagicCookieSize, f);
    if (st) return kALAC_ParamError;
    mConfig.frameLength = f[0];
    mConfig.compatibleVersion = (uint8_t)f[1];
    mConfig.bitDepth = (uint8_t)f[2];
    mConfig.pb = (uint8_t)f[3];
    mConfig.mb = (uint8_t)f[4];
    mConfig.kb = (uint8_t)f[5];
    mConfig.numChannels = (uint8_t)f[6];
    mConfig.maxRun = (uint16_t)f[7];
    mConfig.maxFrameBytes = f[8];
    mConfig.avgBitRate = f[9];
    mConfig.sampleRate = f[10];
    // keep the 24-byte config (without 'frma'/'alac' wrappers) for the batched ABI
    const uint8_t *p = static_cast<const uint8_t *>(inMagicCookie);
    uint32_t left = inMagicCookieSize;
    if (left >= 12 && p[4] == 'f' && p[5] == 'r' && p[6] == 'm' && p[7] == 'a') { p += 12; left -= 12; }
    if (left >= 12 && p[4] == 'a' && p[5] == 'l' && p[6] == 'a' && p[7] == 'c') { p += 12; left -= 12; }
    mCookieSize = 24;
    memcpy(mCookie, p, 24);
    if (!mEngine) {
        st = alac_b200_engine_create(-1, &mEngine);
        if (st != ALAC_B200_OK) return st == ALAC_B200_MEM_ERROR ? kALAC_MemFullError : st;
    }
    return ALAC_noErr;
}

int32_t ALACDecoder::Init(void *inMagicCookie, uint32_t inMagicCookieSize, int /*X*/)
{
    return Init(inMagicCookie, inMagicCookieSize);
}

static uint32_t bytes_per_sample(uint32_t depth) { return depth == 16 ? 2u : depth == 32 ? 4u : 3u; }

int32_t ALACDecoder::Decode(BitBuffer *bits, uint8_t *sampleBuffer, uint32_t numSamples, uint32_t numChannels,
                            uint32_t *outNumSamples)
{
    // codec/ALACDecoder.cu:571-1002 for one packet
    if (!mEngine || !bits || !sampleBuffer || !outNumSamples || numChannels == 0) return kALAC_ParamError;
    (void)numSamples;   // the packet's own header (or the cookie's frameLength) decides, as in the reference
    const uint32_t size = (uint32_t)(bits->end - bits->cur);
    uint64_t frames = 0;
    int32_t status = 0;
    uint32_t n = 0;
    const uint64_t cap = (uint64_t)mConfig.frameLength * mConfig.numChannels * bytes_per_sample(mConfig.bitDepth);
    int32_t st = alac_b200_decode(mEngine, mCookie, mCookieSize, bits->cur, &size, 1, ALAC_B200_MEM_HOST, sampleBuffer, cap,
                                  &n, &status, ALAC_B200_MEM_HOST, &frames, nullptr);
    *outNumSamples = n;
    bits->cur = bits->end;      // the whole packet is consumed
    bits->bitIndex = 0;
    return st;
}

int32_t ALACDecoder::Decode(BitBuffer *bits, uint32_t numSamples, uint32_t numChannels, uint32_t *outNumSamples,
                            uint32_t outBytesPerPacket, int X)
{
    if (X < 0) return kALAC_ParamError;
    if (mStaged.size() <= (size_t)X) mStaged.resize((size_t)X + 1);
    std::vector<uint8_t> &slot = mStaged[(size_t)X];
    const size_t cap = (size_t)mConfig.frameLength * mConfig.numChannels * bytes_per_sample(mConfig.bitDepth);
    slot.assign(cap > outBytesPerPacket ? cap : outBytesPerPacket, 0);
    return Decode(bits, slot.data(), numSamples, numChannels, outNumSamples);
}

extern "C" int32_t alac_b200_copy_to_device(void *dst, const void *src, uint64_t bytes);   // alac_engine.cu

void ALACDecoder::fillWriteBuffer(void *sampleBuffer, uint32_t /*numChannels*/, int32_t theOutputPacketBytes, int X)
{
    // fork: codec/ALACDecoder.cu:497-563 writes every staged packet at X * theOutputPacketBytes of a device buffer
    for (int i = 0; i < X && (size_t)i < mStaged.size(); i++) {
        if (mStaged[(size_t)i].empty()) continue;
        alac_b200_copy_to_device(static_cast<uint8_t *>(sampleBuffer) + (size_t)i * (size_t)theOutputPacketBytes,
                                 mStaged[(size_t)i].data(), (uint64_t)theOutputPacketBytes);
    }
}

int32_t ALACDecoder::DecodeBatch(const unsigned char *packets, const uint32_t *packetSizes, uint64_t numPackets,
                                 unsigned char *pcmOut, uint64_t pcmCap, uint64_t *outSampleFrames, int32_t *packetStatus)
{
    if (!mEngine) return kALAC_ParamError;
    return alac_b200_decode(mEngine, mCookie, mCookieSize, packets, packetSizes, numPackets, ALAC_B200_MEM_HOST, pcmOut,
                            pcmCap, nullptr, packetStatus, ALAC_B200_MEM_HOST, outSampleFrames, nullptr);
}
