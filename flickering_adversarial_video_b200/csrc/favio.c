/* favio.c — see include/favio.h.  gcc -O3 -shared -fPIC -o ../libfavio.so favio.c */
#include "../../include/favio.h"

#include <string.h>

static uint32_t g_tab[8][256];
static int g_init = 0;

static void init_tables(void) {
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
    g_tab[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) g_tab[t][i] = (g_tab[t - 1][i] >> 8) ^ g_tab[0][g_tab[t - 1][i] & 0xffu];
  g_init = 1;
}

uint32_t favio_crc32c(uint32_t crc, const void* data, size_t n) {
  if (!g_init) init_tables();
  const uint8_t* p = (const uint8_t*)data;
  uint32_t c = ~crc;
  while (n && ((uintptr_t)p & 7u)) { c = (c >> 8) ^ g_tab[0][(c ^ *p++) & 0xffu]; --n; }
  while (n >= 8) {   /* slice-by-8 */
    uint64_t v;
    memcpy(&v, p, 8);
    v ^= c;
    c = g_tab[7][v & 0xff] ^ g_tab[6][(v >> 8) & 0xff] ^ g_tab[5][(v >> 16) & 0xff] ^ g_tab[4][(v >> 24) & 0xff] ^
        g_tab[3][(v >> 32) & 0xff] ^ g_tab[2][(v >> 40) & 0xff] ^ g_tab[1][(v >> 48) & 0xff] ^ g_tab[0][v >> 56];
    p += 8;
    n -= 8;
  }
  while (n--) c = (c >> 8) ^ g_tab[0][(c ^ *p++) & 0xffu];
  return ~c;
}

uint32_t favio_masked_crc32c(const void* data, size_t n) {
  const uint32_t c = favio_crc32c(0, data, n);
  return ((c >> 15) | (c << 17)) + 0xa282ead8u;
}

int64_t favio_tfrecord_index(const void* file, size_t n, int verify_payload, uint64_t* offsets, uint64_t* lengths,
                             int64_t cap) {
  const uint8_t* p = (const uint8_t*)file;
  size_t pos = 0;
  int64_t count = 0;
  while (pos < n) {
    uint64_t len;
    uint32_t lcrc, dcrc;
    if (n - pos < 12) return -(1 + count);
    memcpy(&len, p + pos, 8);
    memcpy(&lcrc, p + pos + 8, 4);
    if (favio_masked_crc32c(p + pos, 8) != lcrc) return -(1 + count);
    if (len > n - pos - 12 || n - pos - 12 - len < 4) return -(1 + count);
    if (verify_payload) {
      memcpy(&dcrc, p + pos + 12 + len, 4);
      if (favio_masked_crc32c(p + pos + 12, (size_t)len) != dcrc) return -(1 + count);
    }
    if (count < cap) {
      if (offsets) offsets[count] = pos + 12;
      if (lengths) lengths[count] = len;
    }
    ++count;
    pos += 12 + (size_t)len + 4;
  }
  return count;
}
