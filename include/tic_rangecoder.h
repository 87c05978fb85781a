/*
 * tic_rangecoder.h — C ABI of the host entropy coder that sits right after the hot path
 * (SURVEY.md §8f rank 1).  It stands where the third-party PyPI package `range_coder` stands in the
 * reference: encode.py:76-97 (RangeEncoder(path).encode(list, cum_freq); .close()),
 * decode.py:79-101 (RangeDecoder(path).decode(n, cum_freq)) and prob_to_cum_freq / cum_freq_to_prob
 * (other/test_range_coder.py:186-229).  The package itself is not vendored in the reference and not
 * installable here; this is a restatement of the published carry-propagating byte-wise range coder (the
 * arithmetic and the stream format are in include/tic_rc_core.h, shared with the CUDA entropy stage), pinned by
 * the reference's own known-answer test (other/test_range_coder.py:37-68: 17 x [0,0,0,0,1,2] under cumFreq
 * [0,4,6,8] -> 17 bytes, bytes 4..16 == 0x0b).
 *
 * NOT STREAM-COMPATIBLE WITH THE PyPI PACKAGE: files written by the reference's own range_coder cannot be decoded
 * here and vice versa (that package's first four stream bytes are "special" by its own test's admission; nothing in
 * the reference pins them).  bpp may differ by the coders' termination overhead (a few bytes per file).
 * Frequency totals are limited to 2^16 (the reference uses resolution = 4096, encode.py:91).
 *
 * Symbols may be given as uint8 (the codec's native symbol type: no Python list round trip) or int32.
 * All functions return 0 or a negative tic_rc_status; no exceptions cross the boundary.
 */
#ifndef TIC_RANGECODER_H_
#define TIC_RANGECODER_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tic_rc_encoder tic_rc_encoder;
typedef struct tic_rc_decoder tic_rc_decoder;

typedef enum tic_rc_status {
  TIC_RC_OK = 0,
  TIC_RC_ERR_IO = -1,       /* cannot open / write the file (RuntimeError) */
  TIC_RC_ERR_TABLE = -2,    /* invalid frequency table: empty, too short, not starting at 0, decreasing, total > 2^16 (ValueError) */
  TIC_RC_ERR_SYMBOL = -3,   /* symbol out of range or of zero probability (ValueError) */
  TIC_RC_ERR_CLOSED = -4    /* encode / decode after close (RuntimeError) */
} tic_rc_status;

/* RangeEncoder(filepath) (encode.py:94) */
int tic_rc_encoder_open(tic_rc_encoder** out, const char* path);
/* RangeEncoder.encode(data, cumFreq) (encode.py:95): cum_freq has n_cum = num_symbols + 1 entries, cum_freq[0] == 0 */
int tic_rc_encode_u8(tic_rc_encoder* e, const uint8_t* symbols, int64_t n, const uint32_t* cum_freq, int n_cum);
int tic_rc_encode_i32(tic_rc_encoder* e, const int32_t* symbols, int64_t n, const uint32_t* cum_freq, int n_cum);
/* RangeEncoder.close() (encode.py:97): terminates the stream and closes the file; idempotent */
int tic_rc_encoder_close(tic_rc_encoder* e);
void tic_rc_encoder_free(tic_rc_encoder* e);
/* bytes written so far (after close: the file size) */
int64_t tic_rc_encoder_bytes(const tic_rc_encoder* e);

/* RangeDecoder(filepath) (decode.py:96) */
int tic_rc_decoder_open(tic_rc_decoder** out, const char* path);
/* RangeDecoder.decode(size, cumFreq) (decode.py:97) */
int tic_rc_decode_u8(tic_rc_decoder* d, uint8_t* symbols, int64_t n, const uint32_t* cum_freq, int n_cum);
int tic_rc_decode_i32(tic_rc_decoder* d, int32_t* symbols, int64_t n, const uint32_t* cum_freq, int n_cum);
int tic_rc_decoder_close(tic_rc_decoder* d);
void tic_rc_decoder_free(tic_rc_decoder* d);

/* ---- whole streams in memory, many at once -------------------------------------------------------------
 * The reference codes one image after the other in its Python loop (encode.py:152-202, "To be paralleled"); here
 * the per-image streams of a batch are coded concurrently on a thread pool (n_threads <= 0: all hardware threads).
 * Stream i reads symbols[sym_offsets[i] .. sym_offsets[i+1]) and writes at most out_offsets[i+1] - out_offsets[i]
 * bytes at out + out_offsets[i] (tic_rc_max_encoded_bytes(n) always suffices); out_bytes[i] = stored length.  The
 * bytes are exactly what RangeEncoder(path).encode(stream, cum_freq); close() writes to its file. */
int64_t tic_rc_max_encoded_bytes(int64_t n_symbols);
int tic_rc_encode_streams(const uint8_t* symbols, const int64_t* sym_offsets, int64_t n_streams, const uint32_t* cum_freq,
                          int n_cum, uint8_t* out, const int64_t* out_offsets, int64_t* out_bytes, int n_threads);
/* decode.py:171-208 for a directory of files at once: stream i = in_bytes[i] bytes at in + in_offsets[i]. */
int tic_rc_decode_streams(const uint8_t* in, const int64_t* in_offsets, const int64_t* in_bytes, int64_t n_streams,
                          const uint32_t* cum_freq, int n_cum, uint8_t* symbols, const int64_t* sym_offsets, int n_threads);

/* CRC-32C (Castagnoli, reflected, init/xorout 0xffffffff): the checksum of the TF-V2 checkpoint bundle that
 * utils.restore_params reads (utils/utils.py:84-93); lives in this host-only library so that reading a checkpoint
 * does not need the CUDA library. */
uint32_t tic_rc_crc32c(const void* data, uint64_t n);

/* range_coder.prob_to_cum_freq(prob, resolution) (encode.py:91): cum_freq gets n + 1 entries summing to
 * `resolution`; non-zero probabilities get non-zero width, zero probabilities zero width. */
int tic_rc_prob_to_cum_freq(const double* prob, int n, uint32_t resolution, uint32_t* cum_freq);

#ifdef __cplusplus
}
#endif
#endif /* TIC_RANGECODER_H_ */
