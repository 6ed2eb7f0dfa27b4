// ALACAudioTypes.h -- ABI types of the ALAC class API as re-declared by alac_b200.
// Same names, field order and values as the reference's codec/ALACAudioTypes.h (structs :136-148,
// :162-176; error codes :54-60; limits :68-75; layout tags :115-125) so callers written against
// libalac compile unchanged.  Written from the ABI, not copied.
#ifndef ALACAUDIOTYPES_H
#define ALACAUDIOTYPES_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { kALAC_UnimplementedError = -4, kALAC_FileNotFoundError = -43, kALAC_ParamError = -50, kALAC_MemFullError = -108 };
enum { kALACFormatAppleLossless = 0x616c6163 /* 'alac' */, kALACFormatLinearPCM = 0x6c70636d /* 'lpcm' */ };
enum { kALACMaxChannels = 8, kALACMaxEscapeHeaderBytes = 8, kALACMaxSearches = 16, kALACMaxCoefs = 16,
       kALACDefaultFramesPerPacket = 4096 };
enum { kALACFormatFlagIsFloat = 1, kALACFormatFlagIsBigEndian = 2, kALACFormatFlagIsSignedInteger = 4,
       kALACFormatFlagIsPacked = 8, kALACFormatFlagIsAlignedHigh = 16 };
enum { kALACFormatFlagsNativeEndian = 0 };
enum { kALACCodecFormat = 0x616c6163, kALACVersion = 0, kALACCompatibleVersion = kALACVersion, kALACDefaultFrameSize = 4096 };
#define kChannelAtomSize 12

typedef uint32_t ALACChannelLayoutTag;
typedef double alac_float64_t;
enum {
    kALACChannelLayoutTag_Mono = (100 << 16) | 1, kALACChannelLayoutTag_Stereo = (101 << 16) | 2,
    kALACChannelLayoutTag_MPEG_3_0_B = (113 << 16) | 3, kALACChannelLayoutTag_MPEG_4_0_B = (116 << 16) | 4,
    kALACChannelLayoutTag_MPEG_5_0_D = (120 << 16) | 5, kALACChannelLayoutTag_MPEG_5_1_D = (124 << 16) | 6,
    kALACChannelLayoutTag_AAC_6_1 = (142 << 16) | 7, kALACChannelLayoutTag_MPEG_7_1_B = (127 << 16) | 8
};
static const ALACChannelLayoutTag ALACChannelLayoutTags[kALACMaxChannels] = {
    kALACChannelLayoutTag_Mono, kALACChannelLayoutTag_Stereo, kALACChannelLayoutTag_MPEG_3_0_B,
    kALACChannelLayoutTag_MPEG_4_0_B, kALACChannelLayoutTag_MPEG_5_0_D, kALACChannelLayoutTag_MPEG_5_1_D,
    kALACChannelLayoutTag_AAC_6_1, kALACChannelLayoutTag_MPEG_7_1_B };

typedef struct ALACAudioChannelLayout {
    ALACChannelLayoutTag mChannelLayoutTag;
    uint32_t mChannelBitmap;
    uint32_t mNumberChannelDescriptions;
} ALACAudioChannelLayout;

typedef struct AudioFormatDescription {
    alac_float64_t mSampleRate;
    uint32_t mFormatID;
    uint32_t mFormatFlags;        // encoder side: 1/2/3/4 = 16/20/24/32-bit source (codec/ALACEncoder.cu:1463-1479)
    uint32_t mBytesPerPacket;
    uint32_t mFramesPerPacket;
    uint32_t mBytesPerFrame;
    uint32_t mChannelsPerFrame;
    uint32_t mBitsPerChannel;
    uint32_t mReserved;
} AudioFormatDescription;

// the 24-byte magic cookie body; multi-byte fields are big-endian on the wire
typedef struct ALACSpecificConfig {
    uint32_t frameLength;
    uint8_t compatibleVersion;
    uint8_t bitDepth;
    uint8_t pb;
    uint8_t mb;
    uint8_t kb;
    uint8_t numChannels;
    uint16_t maxRun;
    uint32_t maxFrameBytes;
    uint32_t avgBitRate;
    uint32_t sampleRate;
} ALACSpecificConfig;

enum { AudioChannelLayoutAID = 0x6368616e /* 'chan' */ };

#ifdef __cplusplus
}
#endif
#endif
