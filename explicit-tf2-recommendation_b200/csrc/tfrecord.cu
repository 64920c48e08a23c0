// Input side of the hot path: TFRecord frames of serialized tf.train.Example -> column-major host
// arrays (meant to be PINNED: the Trainer moves back-to-back columns to the device with one async copy per dtype).
//
// Replaces tf.data.TFRecordDataset + tf.io.parse_single_example with FixedLenFeature(shape=[w]) schemas
// (2.FM/ModelManager.py:122-153) for the files 2.FM/DataGenerator.py:104-124 writes:
//   frame   = uint64 length | uint32 masked_crc32c(length) | data[length] | uint32 masked_crc32c(data)
//   Example = { 1: Features { 1: map<string, Feature> } }
//   Feature = { 1: BytesList | 2: FloatList { 1: packed float } | 3: Int64List { 1: packed varint } }
// Host code only (no kernel): it lives in libetr.so so that the parser is native, like the TF op it replaces.
#include <stdint.h>
#include <string.h>

#include "etr_common.cuh"

namespace {

uint32_t g_crc_tab[8][256];
bool g_crc_ready = false;

void crc_init() {
  if (g_crc_ready) return;
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
    g_crc_tab[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) g_crc_tab[t][i] = (g_crc_tab[t - 1][i] >> 8) ^ g_crc_tab[0][g_crc_tab[t - 1][i] & 0xff];
  g_crc_ready = true;
}

uint32_t crc32c(const uint8_t* p, size_t n) {        // slice-by-8
  uint32_t c = 0xffffffffu;
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    w ^= c;
    c = g_crc_tab[7][w & 0xff] ^ g_crc_tab[6][(w >> 8) & 0xff] ^ g_crc_tab[5][(w >> 16) & 0xff] ^ g_crc_tab[4][(w >> 24) & 0xff] ^
        g_crc_tab[3][(w >> 32) & 0xff] ^ g_crc_tab[2][(w >> 40) & 0xff] ^ g_crc_tab[1][(w >> 48) & 0xff] ^ g_crc_tab[0][w >> 56];
    p += 8;
    n -= 8;
  }
  while (n--) c = g_crc_tab[0][(c ^ *p++) & 0xff] ^ (c >> 8);
  return c ^ 0xffffffffu;
}

inline uint32_t masked(uint32_t c) { return ((c >> 15) | (c << 17)) + 0xa282ead8u; }

struct Cur {
  const uint8_t* p;
  const uint8_t* e;
  bool ok;
  uint64_t varint() {
    uint64_t v = 0;
    int sh = 0;
    while (p < e && sh < 64) {
      const uint8_t b = *p++;
      v |= (uint64_t)(b & 0x7f) << sh;
      if (!(b & 0x80)) return v;
      sh += 7;
    }
    ok = false;
    return 0;
  }
  Cur sub() {                       // length-delimited field
    const uint64_t n = varint();
    if (!ok || n > (uint64_t)(e - p)) { ok = false; return Cur{p, p, false}; }
    Cur c{p, p + n, true};
    p += n;
    return c;
  }
  void skip(int wt) {
    if (wt == 0) varint();
    else if (wt == 1) { if (e - p >= 8) p += 8; else ok = false; }
    else if (wt == 2) sub();
    else if (wt == 5) { if (e - p >= 4) p += 4; else ok = false; }
    else ok = false;
  }
};

}  // namespace

extern "C" int etr_tfrecord_parse(const uint8_t* buf, int64_t len, int32_t n_feat, const char* const* names,
                                  const int32_t* kinds, const int32_t* widths, void* const* cols, int64_t capacity,
                                  int32_t verify_crc, int64_t* n_rows, int64_t* consumed) {
  ETR_CHECK_ARG(buf && names && kinds && widths && cols && n_rows && consumed && n_feat > 0 && n_feat <= 4096 && len >= 0,
                "bad argument");
  if (verify_crc) crc_init();
  size_t name_len[4096];
  for (int i = 0; i < n_feat; ++i) {
    ETR_CHECK_ARG(names[i] && cols[i] && widths[i] >= 1 && (kinds[i] == 0 || kinds[i] == 1), "bad feature spec");
    name_len[i] = strlen(names[i]);
  }
  int64_t row = 0, pos = 0;
  int seen_gen[4096];
  for (int i = 0; i < n_feat; ++i) seen_gen[i] = -1;
  while (row < capacity && pos + 12 <= len) {
    uint64_t rec_len;
    memcpy(&rec_len, buf + pos, 8);
    if (rec_len > (uint64_t)(len - pos) || pos + 12 + (int64_t)rec_len + 4 > len) break;     // partial frame: caller refills
    const uint8_t* data = buf + pos + 12;
    if (verify_crc) {
      uint32_t c1, c2;
      memcpy(&c1, buf + pos + 8, 4);
      memcpy(&c2, data + rec_len, 4);
      if (masked(crc32c(buf + pos, 8)) != c1 || masked(crc32c(data, rec_len)) != c2) {
        etr_set_error("etr_tfrecord_parse: checksum mismatch in record %lld (offset %lld)", (long long)row, (long long)pos);
        return ETR_EINVAL;
      }
    }
    Cur ex{data, data + rec_len, true};
    int found = 0;
    while (ex.ok && ex.p < ex.e) {
      const uint64_t tag = ex.varint();
      if ((tag >> 3) != 1 || (tag & 7) != 2) { ex.skip((int)(tag & 7)); continue; }
      Cur feats = ex.sub();                                   // Features
      while (feats.ok && feats.p < feats.e) {
        const uint64_t t2 = feats.varint();
        if ((t2 >> 3) != 1 || (t2 & 7) != 2) { feats.skip((int)(t2 & 7)); continue; }
        Cur entry = feats.sub();                              // map entry: 1 = key, 2 = Feature
        const uint8_t* key = nullptr;
        size_t key_len = 0;
        Cur feat{nullptr, nullptr, false};
        while (entry.ok && entry.p < entry.e) {
          const uint64_t t3 = entry.varint();
          if ((t3 >> 3) == 1 && (t3 & 7) == 2) { Cur k = entry.sub(); key = k.p; key_len = (size_t)(k.e - k.p); }
          else if ((t3 >> 3) == 2 && (t3 & 7) == 2) feat = entry.sub();
          else entry.skip((int)(t3 & 7));
        }
        if (!entry.ok || !key || !feat.ok) { ex.ok = false; break; }
        int fi = -1;
        for (int i = 0; i < n_feat; ++i)
          if (name_len[i] == key_len && memcmp(names[i], key, key_len) == 0) { fi = i; break; }
        if (fi < 0) continue;                                 // a feature the schema does not ask for
        const int w = widths[fi];
        int got = 0;
        while (feat.ok && feat.p < feat.e) {
          const uint64_t t4 = feat.varint();
          const int fno = (int)(t4 >> 3);
          if ((t4 & 7) != 2 || (fno != 2 && fno != 3)) { feat.skip((int)(t4 & 7)); continue; }
          Cur list = feat.sub();                              // FloatList / Int64List
          if ((fno == 3) != (kinds[fi] == 0)) {
            etr_set_error("etr_tfrecord_parse: feature '%s' has the wrong list type in record %lld", names[fi], (long long)row);
            return ETR_EINVAL;
          }
          while (list.ok && list.p < list.e) {
            const uint64_t t5 = list.varint();
            if ((t5 >> 3) != 1) { list.skip((int)(t5 & 7)); continue; }
            if (kinds[fi] == 0) {
              int64_t* dst = (int64_t*)cols[fi] + row * w;
              if ((t5 & 7) == 2) {
                Cur pk = list.sub();
                while (pk.ok && pk.p < pk.e) { const uint64_t v = pk.varint(); if (got < w) dst[got] = (int64_t)v; ++got; }
                if (!pk.ok) list.ok = false;
              } else if ((t5 & 7) == 0) {
                const uint64_t v = list.varint();
                if (got < w) dst[got] = (int64_t)v;
                ++got;
              } else list.ok = false;
            } else {
              float* dst = (float*)cols[fi] + row * w;
              if ((t5 & 7) == 2) {
                Cur pk = list.sub();
                while (pk.ok && pk.e - pk.p >= 4) { if (got < w) memcpy(dst + got, pk.p, 4); pk.p += 4; ++got; }
                if (!pk.ok || pk.p != pk.e) list.ok = false;
              } else if ((t5 & 7) == 5) {
                if (list.e - list.p >= 4) { if (got < w) memcpy(dst + got, list.p, 4); list.p += 4; ++got; } else list.ok = false;
              } else list.ok = false;
            }
          }
          if (!list.ok) feat.ok = false;
        }
        if (!feat.ok) { ex.ok = false; break; }
        if (got != w) {
          etr_set_error("etr_tfrecord_parse: feature '%s' has %d values in record %lld, the schema says %d "
                        "(FixedLenFeature semantics)", names[fi], got, (long long)row, w);
          return ETR_EINVAL;
        }
        if (seen_gen[fi] != (int)(row & 0x7fffffff)) { seen_gen[fi] = (int)(row & 0x7fffffff); ++found; }
      }
      if (!feats.ok) ex.ok = false;
    }
    if (!ex.ok) {
      etr_set_error("etr_tfrecord_parse: malformed tf.train.Example in record %lld (offset %lld)", (long long)row, (long long)pos);
      return ETR_EINVAL;
    }
    if (found != n_feat) {
      for (int i = 0; i < n_feat; ++i)
        if (seen_gen[i] != (int)(row & 0x7fffffff)) {
          etr_set_error("etr_tfrecord_parse: feature '%s' is missing from record %lld (FixedLenFeature without a default)",
                        names[i], (long long)row);
          return ETR_EINVAL;
        }
    }
    pos += 12 + (int64_t)rec_len + 4;
    ++row;
  }
  *n_rows = row;
  *consumed = pos;
  return ETR_OK;
}
