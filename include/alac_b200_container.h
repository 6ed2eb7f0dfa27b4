/*
 * alac_b200_container.h -- CAF / WAV container I/O around the batched codec ABI.
 *
 * Replaces the file plumbing of the reference's alacconvert:
 *   convert-utility/CAFFileALAC.cpp:60-187  WriteCAFF{caff,desc,kuki,chan,pakt,free,data}Chunk
 *   convert-utility/CAFFileALAC.cpp:189-286 GetBERInteger / ReadBERInteger / BuildBasePacketTable
 *   convert-utility/main.cu:200-389         GetInputFormat / FindDataStart (WAV + CAF sniffing)
 *   convert-utility/main.cu:803-852         WriteWAVE{RIFF,fmt,data}Chunk
 * Files written here are byte-identical to what the reference CLI writes for the same packets
 * (layout in SURVEY.md Appendix E, including its quirks).  Host-only code: no GPU work happens here.
 */
#ifndef ALAC_B200_CONTAINER_H
#define ALAC_B200_CONTAINER_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct alac_b200_pcm_info {
    uint32_t sample_rate;
    uint32_t channels;
    uint32_t bit_depth;           /* 16, 20, 24, 32 */
    uint64_t data_offset;         /* byte offset of the first sample in the file */
    uint64_t data_bytes;          /* PCM payload size */
} alac_b200_pcm_info;

typedef struct alac_b200_caf_info {
    uint32_t sample_rate;
    uint32_t channels;
    uint32_t bit_depth;           /* from the desc chunk's format flags 1..4 */
    uint32_t frames_per_packet;
    uint8_t  cookie[64];
    uint32_t cookie_size;
    uint64_t num_packets;         /* entries actually present in the packet table */
    uint64_t valid_frames;        /* pakt header mNumberValidFrames */
    uint64_t table_offset;        /* file offset of the first BER entry */
    uint64_t table_bytes;
    uint64_t data_offset;         /* file offset of the first packet byte (after the 4-byte edit count) */
    uint64_t data_bytes;          /* packet bytes available */
} alac_b200_caf_info;

/* BER integers of the packet table (CAFFileALAC.cpp:189-258). Returns bytes written / consumed (0 on error). */
uint32_t alac_b200_ber_encode(uint32_t value, uint8_t out[5]);
uint32_t alac_b200_ber_decode(const uint8_t *in, uint32_t avail, uint32_t *value);

/* Canonical WAV (RIFF/WAVE, PCM) sniffing: fills info, returns 0 or kALAC_ParamError / kALAC_FileNotFoundError. */
int32_t alac_b200_wav_probe(const char *path, alac_b200_pcm_info *info);
/* 44-byte canonical header + PCM (main.cu:803-852, sizes patched as at :761-769). */
int32_t alac_b200_wav_write(const char *path, uint32_t sample_rate, uint32_t channels, uint32_t bit_depth,
                            const void *pcm, uint64_t pcm_bytes);

/* ALAC-in-CAF exactly as EncodeALAC lays it out (main.cu:418-629): caff, desc, kuki, [chan], pakt
   (worst-case sized table, BER entries), [free], data.  input_pcm_bytes drives the pakt header the way
   BuildBasePacketTable does. */
int32_t alac_b200_caf_write(const char *path, uint32_t sample_rate, uint32_t channels, uint32_t bit_depth,
                            const void *cookie, uint32_t cookie_size, uint64_t input_pcm_bytes,
                            const void *packets, const uint32_t *packet_sizes, uint64_t num_packets);
/* Parse an ALAC CAF: desc, kuki, pakt, data (main.cu:635-734, CAFFileALAC.cpp:288-456). */
int32_t alac_b200_caf_probe(const char *path, alac_b200_caf_info *info);
/* Read the packet table into sizes[] (up to cap entries; stops at the first zero entry like the reference
   decode loop, main.cu:717).  Returns the number of entries. */
uint64_t alac_b200_caf_read_table(const char *path, const alac_b200_caf_info *info, uint32_t *sizes, uint64_t cap);

#ifdef __cplusplus
}
#endif
#endif
