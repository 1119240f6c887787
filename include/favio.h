/* favio — host-side I/O helpers of the "next" rows of SURVEY.md §8 (f1: TFRecord input, f4: TensorBoard scalars).
 * Plain C, no CUDA.  Replaces what the reference gets from TensorFlow's record reader / summary writer:
 *   tf.data.TFRecordDataset            i3d_adversarial_main_universal.py:231-248
 *   tf.python_io.TFRecordWriter        kinetics_to_tf_record_uint8.py:62,95
 *   tf.summary.scalar / SummarySaverHook  i3d_adversarial_main_universal.py:176-201
 * Both file formats frame their payloads as  [u64 length][u32 masked_crc32c(length)][payload][u32 masked_crc32c(payload)]. */
#ifndef FAVIO_H_
#define FAVIO_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* CRC-32C (Castagnoli, reflected polynomial 0x82F63B78), `crc` = running value (0 to start). */
uint32_t favio_crc32c(uint32_t crc, const void* data, size_t n);
/* TensorFlow's masked CRC: rotate right by 15 and add 0xa282ead8 (tensorflow/core/lib/hash/crc32c.h). */
uint32_t favio_masked_crc32c(const void* data, size_t n);
/* Scan a TFRecord file image: writes up to `cap` (payload offset, payload length) pairs, returns the number of
 * records, or -(1 + index) of the first record whose framing or CRCs are wrong.  verify_payload = 0 skips the
 * payload CRC (the length CRC is always checked). */
int64_t favio_tfrecord_index(const void* file, size_t n, int verify_payload, uint64_t* offsets, uint64_t* lengths,
                             int64_t cap);

#ifdef __cplusplus
}
#endif
#endif
