// class_api_demo.cpp -- drives the drop-in ALACEncoder / ALACDecoder exactly the way the
// reference's alacconvert does (convert-utility/main.cu:411-426, :552-601, :707-744): one Encode()
// per 4096-sample frame on a single encoder object, then one Decode() per packet.
// usage: class_api_demo <pcm.raw> <channels> <bitdepth> <samplerate> <out_prefix>
// writes <out_prefix>.cookie, .packets, .sizes (uint32 LE), .pcm
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>
#include "ALACBitUtilities.h"
#include "ALACDecoder.h"
#include "ALACEncoder.h"

static std::vector<uint8_t> slurp(const char *path)
{
    std::vector<uint8_t> v;
    FILE *f = fopen(path, "rb");
    if (!f) return v;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    v.resize((size_t)n);
    if (n && fread(v.data(), 1, (size_t)n, f) != (size_t)n) v.clear();
    fclose(f);
    return v;
}
static void dump(const std::string &path, const void *p, size_t n)
{
    FILE *f = fopen(path.c_str(), "wb");
    if (n) fwrite(p, 1, n, f);
    fclose(f);
}

int main(int argc, char **argv)
{
    if (argc < 6) return 2;
    std::vector<uint8_t> pcm = slurp(argv[1]);
    const uint32_t ch = (uint32_t)atoi(argv[2]), depth = (uint32_t)atoi(argv[3]), rate = (uint32_t)atoi(argv[4]);
    const std::string prefix = argv[5];
    const uint32_t bps = depth == 16 ? 2 : depth == 32 ? 4 : 3;

    AudioFormatDescription in = {}, out = {};
    in.mSampleRate = rate; in.mFormatID = kALACFormatLinearPCM; in.mChannelsPerFrame = ch; in.mBitsPerChannel = depth;
    in.mBytesPerFrame = in.mBytesPerPacket = ch * bps; in.mFramesPerPacket = 1;
    out.mSampleRate = rate; out.mFormatID = kALACFormatAppleLossless; out.mChannelsPerFrame = ch;
    out.mFramesPerPacket = kALACDefaultFramesPerPacket;
    out.mFormatFlags = depth == 16 ? 1 : depth == 20 ? 2 : depth == 24 ? 3 : 4;     // main.cu:285-302

    ALACEncoder enc;
    enc.SetFrameSize(out.mFramesPerPacket);
    if (enc.InitializeEncoder(out) != 0) { fprintf(stderr, "InitializeEncoder failed\n"); return 1; }
    uint8_t cookie[64];
    uint32_t cookieSize = enc.GetMagicCookieSize(ch);
    enc.GetMagicCookie(cookie, &cookieSize);
    dump(prefix + ".cookie", cookie, cookieSize);

    const uint32_t packetBytes = ch * bps * out.mFramesPerPacket;
    std::vector<uint8_t> write(packetBytes + kALACMaxEscapeHeaderBytes * ch + 64), packets;
    std::vector<uint32_t> sizes;
    for (size_t off = 0; off < pcm.size(); off += packetBytes) {
        int32_t n = (int32_t)std::min<size_t>(packetBytes, pcm.size() - off);
        if (enc.Encode(in, out, pcm.data() + off, write.data(), &n) != 0) { fprintf(stderr, "Encode failed\n"); return 1; }
        packets.insert(packets.end(), write.begin(), write.begin() + n);
        sizes.push_back((uint32_t)n);
    }
    enc.Finish();
    dump(prefix + ".packets", packets.data(), packets.size());
    dump(prefix + ".sizes", sizes.data(), sizes.size() * 4);

    ALACDecoder dec;
    if (dec.Init(cookie, cookieSize) != 0) { fprintf(stderr, "Init failed\n"); return 1; }
    std::vector<uint8_t> back, frame(packetBytes);
    size_t at = 0;
    for (uint32_t sz : sizes) {
        BitBuffer bits;
        BitBufferInit(&bits, packets.data() + at, sz);
        uint32_t n = 0;
        if (dec.Decode(&bits, frame.data(), dec.mConfig.frameLength, ch, &n) != 0) { fprintf(stderr, "Decode failed\n"); return 1; }
        back.insert(back.end(), frame.begin(), frame.begin() + (size_t)n * ch * bps);
        at += sz;
    }
    dump(prefix + ".pcm", back.data(), back.size());
    printf("packets=%zu bytes=%zu decoded=%zu\n", sizes.size(), packets.size(), back.size());
    return 0;
}
