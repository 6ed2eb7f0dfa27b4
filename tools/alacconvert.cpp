// alacconvert -- drop-in for the reference's convert-utility (convert-utility/main.cu:73-197):
//   alacconvert <input.wav> <output.caf>     encode PCM WAV -> ALAC in CAF
//   alacconvert <input.caf> <output.wav>     decode ALAC in CAF -> PCM WAV
// Options (extensions): -k N   encoder-reset schedule, frames per segment (default 0 = byte-identical to
//                              the reference CLI; N >= 1 encodes segments in parallel)
//                       -f     fast mode (SetFastMode)
//                       -g N   shard the call by frame range over the first N GPUs of the box
//                              (alac_b200_engine_create_multi; same bytes as one GPU)
// The whole file goes through ONE batched call of libalac_b200; no per-frame loop, no CPU codec.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "alac_b200.h"
#include "alac_b200_container.h"

static bool ends_with(const std::string &s, const char *suf)
{
    const size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

static int usage()
{
    printf("Usage: alacconvert [-k frames_per_segment] [-f] [-g gpus] <input file> <output file>\n"
           "       WAV -> CAF encodes, CAF -> WAV decodes (16/20/24/32-bit PCM, 1-8 channels)\n");
    return 1;
}

int main(int argc, char **argv)
{
    uint32_t K = 0, fast = 0, gpus = 1;
    std::vector<std::string> files;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-h")) return usage();
        else if (!strcmp(argv[i], "-k") && i + 1 < argc) K = (uint32_t)atoi(argv[++i]);
        else if (!strcmp(argv[i], "-f")) fast = 1;
        else if (!strcmp(argv[i], "-g") && i + 1 < argc) gpus = (uint32_t)atoi(argv[++i]);
        else files.push_back(argv[i]);
    }
    if (files.size() != 2) return usage();
    printf("Input file: %s\nOutput file: %s\n", files[0].c_str(), files[1].c_str());   // main.cu:122-123

    alac_b200_engine *eng = nullptr;
    int32_t est;
    if (gpus > 1) {
        std::vector<int32_t> devs;
        for (uint32_t d = 0; d < gpus; d++) devs.push_back((int32_t)d);
        est = alac_b200_engine_create_multi(devs.data(), gpus, &eng);
    } else {
        est = alac_b200_engine_create(-1, &eng);
    }
    if (est != ALAC_B200_OK) { fprintf(stderr, "no usable CUDA device%s\n", gpus > 1 ? "s (or no peer access between them)" : ""); return 1; }
    int rc = 1;
    alac_b200_pcm_info wav;
    alac_b200_caf_info caf;
    if (alac_b200_wav_probe(files[0].c_str(), &wav) == 0) {
        // ---- encode (EncodeALAC, main.cu:391-632)
        std::vector<uint8_t> pcm((size_t)wav.data_bytes);
        FILE *f = fopen(files[0].c_str(), "rb");
        if (!f) { fprintf(stderr, "cannot open %s\n", files[0].c_str()); alac_b200_engine_destroy(eng); return 1; }
        fseek(f, (long)wav.data_offset, SEEK_SET);
        const size_t got = fread(pcm.data(), 1, pcm.size(), f);
        fclose(f);
        const uint64_t bpf = (uint64_t)(wav.bit_depth == 16 ? 2 : wav.bit_depth == 32 ? 4 : 3) * wav.channels;
        const uint64_t frames = got / bpf;
        alac_b200_enc_config cfg = {wav.sample_rate, wav.channels, wav.bit_depth, 4096, fast, K};
        uint8_t cookie[ALAC_B200_COOKIE_MAX];
        const uint32_t cookie_size = alac_b200_magic_cookie(&cfg, 0, 0, cookie, sizeof(cookie));   // main.cu:424-426
        const uint64_t cap = alac_b200_encode_bound(&cfg, frames, 1);
        std::vector<uint8_t> packets((size_t)cap);
        std::vector<uint32_t> sizes((size_t)(frames / 4096 + 2));
        uint64_t np = 0, nb = 0;
        int32_t st = frames ? alac_b200_encode(eng, &cfg, pcm.data(), frames, ALAC_B200_MEM_HOST, nullptr, 1, packets.data(), cap,
                                               sizes.data(), sizes.size(), ALAC_B200_MEM_HOST, nullptr, &np, &nb, nullptr)
                            : 0;
        if (st) fprintf(stderr, "encode failed: %d %s\n", st, alac_b200_last_error(eng));
        else st = alac_b200_caf_write(files[1].c_str(), wav.sample_rate, wav.channels, wav.bit_depth, cookie, cookie_size,
                                      got, packets.data(), sizes.data(), np);
        rc = st ? 1 : 0;
    } else if (alac_b200_caf_probe(files[0].c_str(), &caf) == 0) {
        // ---- decode (DecodeALAC, main.cu:635-778)
        std::vector<uint32_t> sizes((size_t)(caf.table_bytes + 1));
        const uint64_t np = alac_b200_caf_read_table(files[0].c_str(), &caf, sizes.data(), sizes.size());
        uint64_t nbytes = 0;
        for (uint64_t i = 0; i < np; i++) nbytes += sizes[i];
        std::vector<uint8_t> packets((size_t)nbytes);
        FILE *f = fopen(files[0].c_str(), "rb");
        if (!f) { fprintf(stderr, "cannot open %s\n", files[0].c_str()); alac_b200_engine_destroy(eng); return 1; }
        fseek(f, (long)caf.data_offset, SEEK_SET);
        const size_t got = fread(packets.data(), 1, packets.size(), f);
        fclose(f);
        const uint64_t bpf = (uint64_t)(caf.bit_depth == 16 ? 2 : caf.bit_depth == 32 ? 4 : 3) * caf.channels;
        std::vector<uint8_t> pcm((size_t)(np * caf.frames_per_packet * bpf + 16));
        uint64_t frames = 0;
        int32_t st = (got == nbytes && np) ? alac_b200_decode(eng, caf.cookie, caf.cookie_size, packets.data(), sizes.data(), np,
                                                              ALAC_B200_MEM_HOST, pcm.data(), pcm.size(), nullptr, nullptr,
                                                              ALAC_B200_MEM_HOST, &frames, nullptr)
                                           : (np ? -50 : 0);
        if (st) fprintf(stderr, "decode failed: %d %s\n", st, alac_b200_last_error(eng));
        else st = alac_b200_wav_write(files[1].c_str(), caf.sample_rate, caf.channels, caf.bit_depth, pcm.data(), frames * bpf);
        rc = st ? 1 : 0;
    } else {
        fprintf(stderr, "unsupported input: expected PCM WAV (encode) or ALAC CAF (decode)\n");
    }
    alac_b200_engine_destroy(eng);
    return rc;
}
