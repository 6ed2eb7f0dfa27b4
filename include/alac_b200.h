/*
 * alac_b200.h -- C ABI of the B200-native ALAC codec engine (libalac_b200.so).
 *
 * This is the drop-in boundary for the encode/decode hot path of dark-Stallion/alac.
 * The reference has no FFI layer of its own: its boundary is the C++ class API
 *   ALACEncoder  (codec/ALACEncoder.h:34-102)   InitializeEncoder / Encode / GetMagicCookie
 *   ALACDecoder  (codec/ALACDecoder.h:38-72)    Init / Decode
 * driven one 4096-sample frame at a time by convert-utility/main.cu:391-632 (encode loop)
 * and :635-778 (decode loop).  include/ALACEncoder.h and include/ALACDecoder.h re-declare
 * those classes on top of this ABI; the batched entry points below replace the per-frame
 * loops themselves (whole files / streams of frames in one call).
 *
 * Plain pointers and sizes only; every pointer argument is either host memory or CUDA
 * device memory as stated by the accompanying ALAC_B200_MEM_* flag.  All functions return
 * the reference's int32 status codes (codec/ALACAudioTypes.h:54-60):
 *   0 ALAC_noErr, -4 kALAC_UnimplementedError, -50 kALAC_ParamError, -108 kALAC_MemFullError,
 * plus ALAC_B200_CUDA_ERROR for a CUDA runtime failure (text via alac_b200_last_error()).
 * There is NO CPU fallback: every entry point that computes fails with ALAC_B200_CUDA_ERROR
 * when no sm_100 device is usable.
 */
#ifndef ALAC_B200_H
#define ALAC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ALAC_B200_OK            0
#define ALAC_B200_UNIMPLEMENTED (-4)
#define ALAC_B200_PARAM_ERROR   (-50)
#define ALAC_B200_MEM_ERROR     (-108)
#define ALAC_B200_CUDA_ERROR    (-1000)

#define ALAC_B200_MEM_HOST      0
#define ALAC_B200_MEM_DEVICE    1

#define ALAC_B200_MAX_CHANNELS  8
#define ALAC_B200_COOKIE_MAX    48
/* per-stream encoder coefficient state: [channel 8][U,V][row 3, row 7][8 taps] int16
   (the live part of ALACEncoder::mCoefsU/V, codec/ALACEncoder.h:89-90) */
#define ALAC_B200_STATE_INT16S  (8 * 2 * 2 * 8)

typedef struct alac_b200_engine alac_b200_engine;

/* Encoder configuration.  Mirrors what InitializeEncoder()/SetFrameSize()/SetFastMode() take
   (codec/ALACEncoder.cu:1457-1535, codec/ALACEncoder.h:44-47). */
typedef struct alac_b200_enc_config {
    uint32_t sample_rate;
    uint32_t channels;            /* 1..8; element layout follows sChannelMaps, codec/ALACEncoder.cu:97-107 */
    uint32_t bit_depth;           /* 16, 20, 24, 32 (mFormatFlags 1..4, codec/ALACEncoder.cu:1463-1479) */
    uint32_t frame_size;          /* samples per packet, kALACDefaultFrameSize = 4096; <= 16384 */
    uint32_t fast_mode;           /* SetFastMode(): EncodeStereoFast path, codec/ALACEncoder.cu:564-743 */
    /* Encoder-reset schedule.  The output equals what libalac produces when a fresh
       ALACEncoder is started at every frames_per_segment-th frame of each stream.  0 = never reset:
       byte-identical to one ALACEncoder fed the whole stream (serial per stream). */
    uint32_t frames_per_segment;
} alac_b200_enc_config;

/* One input stream inside a batch: a run of interleaved PCM sample-frames. */
typedef struct alac_b200_stream {
    uint64_t first_sample_frame;  /* offset into the pcm buffer, in sample-frames */
    uint64_t num_sample_frames;   /* length; the last packet of a stream may be partial */
} alac_b200_stream;

typedef struct alac_b200_stats {
    uint64_t num_packets;
    uint64_t payload_bytes;       /* sum of packet sizes */
    uint64_t escape_elements;     /* elements written uncompressed */
    uint32_t max_packet_bytes;
    uint32_t kernel_launches;     /* CUDA kernels launched by the call */
    float    ms_h2d, ms_kernels, ms_d2h;   /* CUDA-event times of the three phases */
    float    ms_search;           /* encode: sum over launches of the search kernels (stages A+B; A+B+C when not split) */
    float    ms_assemble;         /* encode: sum over launches of enc_assemble_kernel */
    float    ms_decode;           /* decode: dec_fused_kernel + dec_entropy_kernel + dec_finish_kernel */
    float    ms_final;            /* encode: sum over launches of enc_final_kernel (stage C, split form only) */
    float    ms_entropy;          /* decode: sum over launches of dec_entropy_kernel (groups the fused kernel does not take) */
    float    ms_finish;           /* decode: sum over launches of dec_finish_kernel (same) */
    float    ms_fused;            /* decode: sum over launches of dec_fused_kernel (regular mono / stereo groups) */
    uint32_t final_form;          /* encode, split form: 0 = enc_final_kernel, 1 = enc_final2_kernel one-warp, 2 = two-warp (last launch) */
    uint32_t search_dense;        /* encode, split form: 1 = the search passes streamed PCM through the block ring */
} alac_b200_stats;

/* ---- engine ------------------------------------------------------------------------------ */
/* device < 0 selects the current CUDA device.  The engine owns its scratch and one stream. */
int32_t     alac_b200_engine_create(int32_t device, alac_b200_engine **out_engine);
void        alac_b200_engine_destroy(alac_b200_engine *engine);
const char *alac_b200_last_error(const alac_b200_engine *engine);
const char *alac_b200_version(void);
/* make later calls run on `cuda_stream` (a cudaStream_t) instead of the engine's own non-blocking stream; NULL
   selects the engine's own stream again.  Device buffers must be ready on the stream the call runs on: pass the
   producer's stream (cudaStreamLegacy for work issued on the legacy default stream). */
int32_t     alac_b200_engine_set_stream(alac_b200_engine *engine, void *cuda_stream);

/* ---- several GPUs of one box (no reference analogue: the fork is single-GPU) ---------------
 * One engine that shards every encode / decode call by frame range over `n_devices` GPUs of this process
 * (peer access is enabled between them; devices[0] is the "home" device).  Encode shards whole segments
 * (frames_per_segment), decode shards packets; the result is byte-identical to a single-GPU call.
 * With device output the buffers must live on the home device: every GPU writes its packets straight to their final
 * offset there over NVLink (packet-offset scan + peer stores inside enc_assemble_kernel, no collective, no staging).
 * With host buffers every GPU moves its own range over its own PCIe link.
 * frames_per_segment = 0 cannot be split inside a stream: whole streams are dealt out instead. */
int32_t     alac_b200_engine_create_multi(const int32_t *devices, uint32_t n_devices, alac_b200_engine **out_engine);
uint32_t    alac_b200_engine_num_devices(const alac_b200_engine *engine);

/* ---- magic cookie: ALACEncoder::GetMagicCookie, codec/ALACEncoder.cu:1109-1140 ------------- */
/* Returns the cookie size (24, or 48 for > 2 channels) or 0 when cap is too small. */
uint32_t    alac_b200_magic_cookie(const alac_b200_enc_config *cfg, uint32_t max_frame_bytes,
                                   uint32_t avg_bit_rate, void *out_cookie, uint32_t cap);
/* worst-case bytes alac_b200_encode can write for the given input size */
uint64_t    alac_b200_encode_bound(const alac_b200_enc_config *cfg, uint64_t num_sample_frames, uint64_t num_streams);

/* ---- batched encode: replaces the Encode() loop of convert-utility/main.cu:552-601 --------- */
/*
 * pcm            interleaved little-endian packed PCM (16-bit: int16; 20/24-bit: 3 bytes,
 *                20-bit left-justified; 32-bit: int32), host or device (pcm_mem).
 * streams        n_streams descriptors (host memory); NULL means one stream covering
 *                [0, num_sample_frames).  Streams must not overlap.
 * packets_out    packets back to back in stream order, packet_sizes[i] bytes each
 *                (the 'data' chunk payload and the pakt entries of SURVEY.md App. E);
 *                capacity >= alac_b200_encode_bound().  out_mem says where both outputs live.
 * coef_state     optional, host memory, n_streams * ALAC_B200_STATE_INT16S (= 256) int16: when non-NULL the
 *                first segment of each stream starts from this state instead of init_coefs and the
 *                state after the last frame is written back (incremental streaming / Encode()).
 */
int32_t alac_b200_encode(alac_b200_engine *engine, const alac_b200_enc_config *cfg,
                         const void *pcm, uint64_t num_sample_frames, int32_t pcm_mem,
                         const alac_b200_stream *streams, uint64_t n_streams,
                         void *packets_out, uint64_t packets_cap,
                         uint32_t *packet_sizes, uint64_t sizes_cap, int32_t out_mem,
                         int16_t *coef_state,
                         uint64_t *out_num_packets, uint64_t *out_bytes,
                         alac_b200_stats *stats);

/* ---- encode one rank's frame range of a job that several GPUs share (one engine / process per GPU) -------------
 * The cross-GPU step of SURVEY 8e inside the call: every rank scans its packet sizes, publishes its byte total in
 * `exchange` (a zeroed ALAC_B200_EXCHANGE_BYTES block in the DESTINATION GPU's memory), waits on the device for the
 * totals of the ranks in front of it, and its assemble kernel stores every packet at its final offset of
 * `dst_packets` -- peer stores over NVLink when the destination is another GPU (a pointer obtained with
 * cudaDeviceEnablePeerAccess in one process, or alac_b200_ipc_open across processes).  No collective, no staging copy.
 * The rank whose `rank` is `home_rank` returns only after every rank's block is in place.
 * All ranks pass the same `epoch`, incremented from call to call (it tags the exchange slots). */
#define ALAC_B200_EXCHANGE_BYTES 1024
#define ALAC_B200_MAX_RANKS      16
typedef struct alac_b200_placement {
    void     *dst_packets;     /* device pointer, local or peer: the job's single packet buffer */
    uint64_t  dst_capacity;    /* bytes */
    uint32_t *dst_sizes;       /* device pointer, local or peer: the job's packet size table; may be NULL */
    uint64_t  first_packet;    /* index of this rank's first packet in that table */
    void     *exchange;        /* device pointer, local or peer: ALAC_B200_EXCHANGE_BYTES, zeroed once */
    uint32_t  rank, n_ranks, home_rank;
    uint32_t  epoch;           /* > 0 */
    /* Staged form (staging != NULL; home_rank must be 0).  A rank's final offset depends on every byte of the ranks in
       front of it, so direct placement can only start when all ranks have finished computing.  Staged placement moves the
       bytes DURING the computation instead: the call runs as a pipeline of chunks, each chunk's packets leave for the
       rank's reserved slot staging + slot_offsets[rank] over NVLink (asynchronous peer copies) while the next chunk's
       kernels run, and when every rank is done the home rank closes the gaps between the slots with copies inside its
       own memory.  slot_offsets: n_ranks entries (host memory), 16-byte aligned, slot r at least as large as rank r's encode
       bound; the staging area extends at least 32 bytes past the last slot. */
    void           *staging;        /* device pointer in the home GPU's memory (local or peer), or NULL = direct form */
    const uint64_t *slot_offsets;
    /* staged form: 0 = the home rank's call returns when the job's buffer is complete, another rank's call when its
       block has arrived in its slot; 1 = the call returns with this rank's own kernels.  On the home rank "wait for
       every rank, close the gaps" then keeps running on the device, on the other ranks the block keeps travelling over
       NVLink (*out_base is not known yet and reads 0) -- the caller does other work on this GPU (decoding its own
       shard out of *out_local_block) and ends the job with alac_b200_placed_finish() on EVERY rank. */
    uint32_t        defer_finish;
} alac_b200_placement;
/* Same arguments as alac_b200_encode for this rank's PCM; packet_sizes[] (the rank's own entries) stays local
   (out_mem says where); *out_base receives the byte offset of the rank's block inside dst_packets.  *out_local_block
   (optional) receives a device pointer to the rank's own copy of its block (staged form: engine scratch, valid until the
   engine's next encode call; direct form and the home rank: the block's place inside dst_packets). */
int32_t alac_b200_encode_placed(alac_b200_engine *engine, const alac_b200_enc_config *cfg,
                                const void *pcm, uint64_t num_sample_frames, int32_t pcm_mem,
                                const alac_b200_stream *streams, uint64_t n_streams,
                                const alac_b200_placement *placement,
                                uint32_t *packet_sizes, uint64_t sizes_cap, int32_t out_mem,
                                uint64_t *out_num_packets, uint64_t *out_bytes, uint64_t *out_base,
                                void **out_local_block, alac_b200_stats *stats);

/* ends a staged job whose call was made with defer_finish = 1.  Home rank: returns when every rank's block is at its final
   offset of dst_packets; *out_job_bytes (optional) = the job's total packet bytes.  Other ranks: returns when this rank's
   block has arrived on the home GPU.  A no-op (status 0) when nothing is pending.
   alac_b200_placed_base: the byte offset of this rank's block inside dst_packets of the engine's latest placed call
   (what *out_base of a non-deferred call returns); valid once nothing is pending. */
int32_t alac_b200_placed_finish(alac_b200_engine *engine, uint64_t *out_job_bytes);
int32_t alac_b200_placed_base(alac_b200_engine *engine, uint64_t *out_base);

/* device memory that can be shared with other processes (plain cudaMalloc on the engine's device) and the
   cudaIpc* wrappers a one-process-per-GPU launcher needs; handle = 64 bytes (cudaIpcMemHandle_t) */
int32_t alac_b200_device_alloc(alac_b200_engine *engine, uint64_t bytes, void **out_ptr);
int32_t alac_b200_device_free(alac_b200_engine *engine, void *ptr);
int32_t alac_b200_ipc_export(alac_b200_engine *engine, void *ptr, void *out_handle64);
int32_t alac_b200_ipc_open(alac_b200_engine *engine, const void *handle64, void **out_ptr);
int32_t alac_b200_ipc_close(alac_b200_engine *engine, void *ptr);

/* ---- batched decode: replaces the Decode() loop of convert-utility/main.cu:717-744 --------- */
/*
 * cookie         magic cookie as stored in the CAF 'kuki' chunk (ALACDecoder::Init,
 *                codec/ALACDecoder.cu:109-190; 'frma'/'alac' wrappers are skipped).
 * packets        packets back to back, packet_sizes[i] bytes each (in_mem says where both live).
 * pcm_out        decoded interleaved PCM, packets contiguous in order; capacity in bytes.
 * packet_samples optional (same memory space as pcm_out): sample-frames produced by each packet.
 * packet_status  optional (same memory space as pcm_out): per-packet status (0 or kALAC_ParamError).
 * Returns 0, or the first non-zero packet status.
 */
int32_t alac_b200_decode(alac_b200_engine *engine, const void *cookie, uint32_t cookie_size,
                         const void *packets, const uint32_t *packet_sizes, uint64_t num_packets,
                         int32_t in_mem,
                         void *pcm_out, uint64_t pcm_cap,
                         uint32_t *packet_samples, int32_t *packet_status, int32_t out_mem,
                         uint64_t *out_sample_frames,
                         alac_b200_stats *stats);

/* ---- asynchronous forms ------------------------------------------------------------------------------------------
 * Same arguments as alac_b200_encode / alac_b200_decode; the call runs on a worker thread of the engine and
 * alac_b200_wait returns its status (every output, *out_* word and stats struct is valid only after the wait).
 * One call per engine may be in flight; the config, stream list and cookie are copied at submit, the data buffers must
 * stay valid until the wait.  Two engines on one GPU thereby overlap an encode with a decode: the PCM going up for the
 * next encode shares the PCIe link with the PCM coming down from the previous decode (the link is full duplex), which a
 * strictly alternating caller of the synchronous forms cannot do. */
int32_t alac_b200_encode_submit(alac_b200_engine *engine, const alac_b200_enc_config *cfg,
                                const void *pcm, uint64_t num_sample_frames, int32_t pcm_mem,
                                const alac_b200_stream *streams, uint64_t n_streams,
                                void *packets_out, uint64_t packets_cap,
                                uint32_t *packet_sizes, uint64_t sizes_cap, int32_t out_mem,
                                int16_t *coef_state,
                                uint64_t *out_num_packets, uint64_t *out_bytes,
                                alac_b200_stats *stats);
int32_t alac_b200_decode_submit(alac_b200_engine *engine, const void *cookie, uint32_t cookie_size,
                                const void *packets, const uint32_t *packet_sizes, uint64_t num_packets,
                                int32_t in_mem,
                                void *pcm_out, uint64_t pcm_cap,
                                uint32_t *packet_samples, int32_t *packet_status, int32_t out_mem,
                                uint64_t *out_sample_frames,
                                alac_b200_stats *stats);
int32_t alac_b200_wait(alac_b200_engine *engine);

/* ---- CAF packet table on the device ------------------------------------------------ */
/*
 * Turns the BER-coded size table of a CAF 'pakt' chunk (convert-utility/CAFFileALAC.cpp:189-258) into
 * packet_sizes[] with GPU kernels (mark entry ends, scan, assemble), stopping where the reference's decode
 * loop stops (convert-utility/main.cu:717: a zero size, or a packet that no longer fits data_bytes).
 * The result can be passed straight to alac_b200_decode(..., in_mem = ALAC_B200_MEM_DEVICE).
 */
int32_t alac_b200_ber_table_sizes(alac_b200_engine *engine, const void *table, uint64_t table_bytes, int32_t table_mem,
                                  uint64_t data_bytes, uint32_t *sizes_out, uint64_t sizes_cap, int32_t out_mem,
                                  uint64_t *out_num_packets);

/* The other direction (convert-utility/CAFFileALAC.cpp:189-222): packet_sizes[] -> the BER bytes of a 'pakt' chunk
 * (without its 24-byte header), built on the device (entry lengths, scan, emit).  With device buffers the table of a
 * rank's packet block can be sent to its place in the file image with the block itself. */
int32_t alac_b200_ber_table_build(alac_b200_engine *engine, const uint32_t *packet_sizes, uint64_t num_packets, int32_t sizes_mem,
                                  void *table_out, uint64_t table_cap, int32_t out_mem, uint64_t *out_table_bytes);

/* parse a cookie on the host (no GPU work): fills the 11 ALACSpecificConfig fields in order
   frameLength, compatibleVersion, bitDepth, pb, mb, kb, numChannels, maxRun, maxFrameBytes,
   avgBitRate, sampleRate (codec/ALACAudioTypes.h:162-176) */
int32_t alac_b200_parse_cookie(const void *cookie, uint32_t cookie_size, uint32_t out_fields[11]);

#ifdef __cplusplus
}
#endif
#endif /* ALAC_B200_H */
