// alac_container.cpp -- CAF / WAV container I/O (include/alac_b200_container.h).  Host-only.
#include "../../include/alac_b200_container.h"

#include <cstdio>
#include <cstring>
#include <vector>

namespace {

const int32_t kParamError = -50, kFileNotFound = -43;

struct File {
    FILE *f;
    explicit File(const char *path, const char *mode) : f(fopen(path, mode)) {}
    ~File() { if (f) fclose(f); }
};

void put_be32(std::vector<uint8_t> &v, uint32_t x) { for (int s = 24; s >= 0; s -= 8) v.push_back((uint8_t)(x >> s)); }
void put_be64(std::vector<uint8_t> &v, uint64_t x) { for (int s = 56; s >= 0; s -= 8) v.push_back((uint8_t)(x >> s)); }
void put_tag(std::vector<uint8_t> &v, const char *t) { v.insert(v.end(), t, t + 4); }
uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
uint64_t be64(const uint8_t *p) { return ((uint64_t)be32(p) << 32) | be32(p + 4); }
uint32_t le32(const uint8_t *p) { return ((uint32_t)p[3] << 24) | ((uint32_t)p[2] << 16) | ((uint32_t)p[1] << 8) | p[0]; }
uint32_t bytes_per_sample(uint32_t depth) { return depth == 16 ? 2u : depth == 32 ? 4u : 3u; }

// ALACChannelLayoutTags, codec/ALACAudioTypes.h:115-125
const uint32_t kLayoutTags[8] = {(100u << 16) | 1, (101u << 16) | 2, (113u << 16) | 3, (116u << 16) | 4,
                                 (120u << 16) | 5, (124u << 16) | 6, (142u << 16) | 7, (127u << 16) | 8};

}  // namespace

extern "C" {

uint32_t alac_b200_ber_encode(uint32_t value, uint8_t out[5])
{
    // CAFFileALAC.cpp:189-236: 7 bits per byte, most significant group first, high bit = "more"
    uint32_t n = value < (1u << 7) ? 1 : value < (1u << 14) ? 2 : value < (1u << 21) ? 3 : value < (1u << 28) ? 4 : 5;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t shift = 7 * (n - 1 - i);
        out[i] = (uint8_t)(((value >> shift) & 0x7f) | (i + 1 < n ? 0x80 : 0));
    }
    return n;
}

uint32_t alac_b200_ber_decode(const uint8_t *in, uint32_t avail, uint32_t *value)
{
    // CAFFileALAC.cpp:238-258
    uint32_t v = 0;
    for (uint32_t i = 0; i < avail && i < 5; i++) {
        v = (v << 7) | (in[i] & 0x7f);
        if (!(in[i] & 0x80)) { *value = v; return i + 1; }
    }
    *value = 0;
    return 0;
}

int32_t alac_b200_wav_probe(const char *path, alac_b200_pcm_info *info)
{
    // main.cu:200-275 (fmt) and :345-389 (data)
    if (!path || !info) return kParamError;
    memset(info, 0, sizeof(*info));
    File in(path, "rb");
    if (!in.f) return kFileNotFound;
    uint8_t h[12];
    if (fread(h, 1, 12, in.f) != 12 || memcmp(h, "RIFF", 4) || memcmp(h + 8, "WAVE", 4)) return kParamError;
    bool have_fmt = false;
    for (;;) {
        uint8_t ch[8];
        if (fread(ch, 1, 8, in.f) != 8) break;
        const uint32_t size = le32(ch + 4);
        if (!memcmp(ch, "fmt ", 4)) {
            uint8_t f[16];
            if (size < 16 || fread(f, 1, 16, in.f) != 16) return kParamError;
            const uint32_t code = f[0] | (f[1] << 8);
            if (code != 1 && code != 0xFFFE) return kParamError;       // PCM only
            info->channels = f[2] | (f[3] << 8);
            info->sample_rate = le32(f + 4);
            info->bit_depth = f[14] | (f[15] << 8);
            have_fmt = true;
            fseek(in.f, (long)(size - 16 + (size & 1)), SEEK_CUR);
        } else if (!memcmp(ch, "data", 4)) {
            info->data_offset = (uint64_t)ftell(in.f);
            info->data_bytes = size;
            break;
        } else {
            fseek(in.f, (long)(size + (size & 1)), SEEK_CUR);
        }
    }
    if (!have_fmt || info->data_offset == 0) return kParamError;
    if (!(info->bit_depth == 16 || info->bit_depth == 20 || info->bit_depth == 24 || info->bit_depth == 32)) return kParamError;
    if (info->channels < 1 || info->channels > 8) return kParamError;
    return 0;
}

int32_t alac_b200_wav_write(const char *path, uint32_t sample_rate, uint32_t channels, uint32_t bit_depth,
                            const void *pcm, uint64_t pcm_bytes)
{
    // main.cu:803-852; sizes as patched at :761-769 (RIFF size = data + 'WAVE' + data header + fmt chunk)
    if (!path || (!pcm && pcm_bytes)) return kParamError;
    File out(path, "wb");
    if (!out.f) return kFileNotFound;
    const uint32_t block = channels * bytes_per_sample(bit_depth);
    const uint32_t rate_bytes = sample_rate * block;
    const uint32_t data = (uint32_t)pcm_bytes, riff = (uint32_t)(pcm_bytes + 4 + 8 + 24);
    uint8_t h[44] = {'R', 'I', 'F', 'F', 0, 0, 0, 0, 'W', 'A', 'V', 'E', 'f', 'm', 't', ' ', 16, 0, 0, 0, 1, 0};
    h[4] = (uint8_t)riff; h[5] = (uint8_t)(riff >> 8); h[6] = (uint8_t)(riff >> 16); h[7] = (uint8_t)(riff >> 24);
    h[22] = (uint8_t)channels;
    h[24] = (uint8_t)sample_rate; h[25] = (uint8_t)(sample_rate >> 8); h[26] = (uint8_t)(sample_rate >> 16); h[27] = (uint8_t)(sample_rate >> 24);
    h[28] = (uint8_t)rate_bytes; h[29] = (uint8_t)(rate_bytes >> 8); h[30] = (uint8_t)(rate_bytes >> 16); h[31] = (uint8_t)(rate_bytes >> 24);
    h[32] = (uint8_t)block;
    h[34] = (uint8_t)bit_depth;
    memcpy(h + 36, "data", 4);
    h[40] = (uint8_t)data; h[41] = (uint8_t)(data >> 8); h[42] = (uint8_t)(data >> 16); h[43] = (uint8_t)(data >> 24);
    if (fwrite(h, 1, 44, out.f) != 44) return kParamError;
    if (pcm_bytes && fwrite(pcm, 1, (size_t)pcm_bytes, out.f) != pcm_bytes) return kParamError;
    return 0;
}

int32_t alac_b200_caf_write(const char *path, uint32_t sample_rate, uint32_t channels, uint32_t bit_depth,
                            const void *cookie, uint32_t cookie_size, uint64_t input_pcm_bytes,
                            const void *packets, const uint32_t *packet_sizes, uint64_t num_packets)
{
    if (!path || !cookie || (num_packets && (!packets || !packet_sizes))) return kParamError;
    if (channels < 1 || channels > 8) return kParamError;
    const uint32_t flags = bit_depth == 16 ? 1u : bit_depth == 20 ? 2u : bit_depth == 24 ? 3u : bit_depth == 32 ? 4u : 0u;
    if (!flags) return kParamError;
    // the 'desc' and 'pakt' arithmetic below is the reference's: 4096 frames per packet (kALACDefaultFramesPerPacket,
    // convert-utility/main.cu:286, CAFFileALAC.cpp:260-286).  A cookie that says otherwise would give a file whose chunks disagree.
    if (cookie_size < 24) return kParamError;
    {
        const uint8_t *c = static_cast<const uint8_t *>(cookie);
        if (((uint32_t)c[0] << 24 | (uint32_t)c[1] << 16 | (uint32_t)c[2] << 8 | c[3]) != 4096u) return kParamError;
    }
    File out(path, "wb");
    if (!out.f) return kFileNotFound;

    std::vector<uint8_t> h;
    // 'caff' file header, CAFFileALAC.cpp:60-65
    put_tag(h, "caff"); h.push_back(0); h.push_back(1); h.push_back(0); h.push_back(0);
    // 'desc', :67-96 with the output format of main.cu:277-309 (VBR: bytes/packet = bits/channel = 0)
    put_tag(h, "desc"); put_be64(h, 32);
    {
        double sr = (double)sample_rate;
        uint64_t bits;
        memcpy(&bits, &sr, 8);
        put_be64(h, bits);
        put_tag(h, "alac"); put_be32(h, flags); put_be32(h, 0); put_be32(h, 4096); put_be32(h, channels); put_be32(h, 0);
    }
    // 'kuki', :105-112
    put_tag(h, "kuki"); put_be64(h, cookie_size);
    h.insert(h.end(), static_cast<const uint8_t *>(cookie), static_cast<const uint8_t *>(cookie) + cookie_size);
    // 'chan' for > 2 channels, :129-140
    if (channels > 2) { put_tag(h, "chan"); put_be64(h, 12); put_be32(h, kLayoutTags[channels - 1]); put_be32(h, 0); put_be32(h, 0); }

    // 'pakt', BuildBasePacketTable :260-286.  The quirk is kept: when the length is an exact multiple of 4096
    // the header counts one packet too many and says remainder = 4096.
    const uint64_t bpf = (uint64_t)bytes_per_sample(bit_depth) * channels;
    const uint64_t valid_frames = input_pcm_bytes / bpf;
    uint64_t hdr_packets = valid_frames / 4096;
    const uint32_t remainder = 4096 - (uint32_t)(valid_frames - hdr_packets * 4096);
    if (remainder) hdr_packets += 1;
    const uint32_t entry = (bpf * 4096 + 8 < 16384) ? 2u : 3u;
    uint64_t table_size = (uint64_t)entry * hdr_packets;
    std::vector<uint8_t> table;
    uint64_t data_bytes = 0;
    for (uint64_t i = 0; i < num_packets; i++) {
        uint8_t b[5];
        const uint32_t n = alac_b200_ber_encode(packet_sizes[i], b);
        table.insert(table.end(), b, b + n);
        data_bytes += packet_sizes[i];
    }
    if (table.size() > table_size) table_size = table.size();           // cannot happen for sizes from this encoder
    const uint64_t left = table_size - table.size();
    const bool free_chunk = left > 12;                                   // main.cu:613-622
    put_tag(h, "pakt"); put_be64(h, (free_chunk ? table.size() : table_size) + 24);
    put_be64(h, hdr_packets); put_be64(h, valid_frames); put_be32(h, 0); put_be32(h, remainder);
    h.insert(h.end(), table.begin(), table.end());
    if (free_chunk) { put_tag(h, "free"); put_be64(h, left - 12); h.insert(h.end(), (size_t)(left - 12), 0); }   // :142-161
    else h.insert(h.end(), (size_t)left, 0);
    // 'data', :98-103: size covers the 4-byte edit count (= 1) + packets
    put_tag(h, "data"); put_be64(h, data_bytes + 4); put_be32(h, 1);
    if (fwrite(h.data(), 1, h.size(), out.f) != h.size()) return kParamError;
    if (data_bytes && fwrite(packets, 1, (size_t)data_bytes, out.f) != data_bytes) return kParamError;
    return 0;
}

int32_t alac_b200_caf_probe(const char *path, alac_b200_caf_info *info)
{
    if (!path || !info) return kParamError;
    memset(info, 0, sizeof(*info));
    File in(path, "rb");
    if (!in.f) return kFileNotFound;
    uint8_t h[8];
    if (fread(h, 1, 8, in.f) != 8 || memcmp(h, "caff", 4)) return kParamError;
    fseek(in.f, 0, SEEK_END);
    const uint64_t file_size = (uint64_t)ftell(in.f);
    fseek(in.f, 8, SEEK_SET);
    bool have_desc = false, have_pakt = false, have_data = false;
    for (;;) {
        uint8_t ch[12];
        if (fread(ch, 1, 12, in.f) != 12) break;
        uint64_t size = be64(ch + 4);
        const uint64_t body = (uint64_t)ftell(in.f);
        // chunk sizes come from the file: every chunk but a to-end-of-file 'data' must lie inside it (a negative or
        // huge size would otherwise seek backwards and walk the same chunks forever)
        if (memcmp(ch, "data", 4) && (size > file_size || body + size > file_size)) return kParamError;
        if (!memcmp(ch, "desc", 4)) {
            uint8_t d[32];
            if (size < 32 || fread(d, 1, 32, in.f) != 32) return kParamError;
            uint64_t bits = be64(d);
            double sr;
            memcpy(&sr, &bits, 8);
            info->sample_rate = (uint32_t)sr;
            if (memcmp(d + 8, "alac", 4)) return kParamError;           // decode side takes ALAC only
            const uint32_t flags = be32(d + 12);
            info->bit_depth = flags == 1 ? 16 : flags == 2 ? 20 : flags == 3 ? 24 : flags == 4 ? 32 : 0;
            info->frames_per_packet = be32(d + 20);
            info->channels = be32(d + 24);
            have_desc = true;
        } else if (!memcmp(ch, "kuki", 4)) {
            if (size > sizeof(info->cookie)) return kParamError;
            if (fread(info->cookie, 1, (size_t)size, in.f) != size) return kParamError;
            info->cookie_size = (uint32_t)size;
        } else if (!memcmp(ch, "pakt", 4)) {
            uint8_t p[24];
            if (size < 24 || fread(p, 1, 24, in.f) != 24) return kParamError;
            info->valid_frames = be64(p + 8);
            info->table_offset = body + 24;
            info->table_bytes = size - 24;
            have_pakt = true;
        } else if (!memcmp(ch, "data", 4)) {
            if (size == (uint64_t)-1 || body + size > file_size) size = file_size - body;     // "to end of file"
            if (size < 4) return kParamError;
            info->data_offset = body + 4;
            info->data_bytes = size - 4;
            have_data = true;
            break;                                                       // data is the last chunk the writer emits
        }
        fseek(in.f, (long)(body + size), SEEK_SET);
    }
    if (!have_desc || !have_pakt || !have_data || !info->cookie_size || !info->bit_depth) return kParamError;
    return 0;
}

uint64_t alac_b200_caf_read_table(const char *path, const alac_b200_caf_info *info, uint32_t *sizes, uint64_t cap)
{
    if (!path || !info || !sizes) return 0;
    File in(path, "rb");
    if (!in.f) return 0;
    fseek(in.f, 0, SEEK_END);
    const uint64_t file_size = (uint64_t)ftell(in.f);
    if (info->table_offset > file_size || info->table_bytes > file_size - info->table_offset) return 0;     // (a caller-made info)
    std::vector<uint8_t> t((size_t)info->table_bytes);
    fseek(in.f, (long)info->table_offset, SEEK_SET);
    if (fread(t.data(), 1, t.size(), in.f) != t.size()) return 0;
    uint64_t n = 0, at = 0, used = 0;
    while (n < cap && at < t.size()) {
        uint32_t v = 0;
        const uint32_t k = alac_b200_ber_decode(t.data() + at, (uint32_t)(t.size() - at), &v);
        if (!k || v == 0 || used + v > info->data_bytes) break;          // main.cu:717 stops at a zero size / short read
        sizes[n++] = v;
        at += k;
        used += v;
    }
    return n;
}

}  // extern "C"
