// io_tagged.cuh -- the FFaTag header shared by every FEDEM binary file (.fmx, .fsm, .frs): 30-character
// tag, 16-bit endian mark 0x1234, 8-byte checksum field, ";1.0;\n" (fedem-foundation/src/FFaLib/FFaOS/
// FFaTag.C:192-297; src/vpmUtilities/binaryDB.c:643-733).  Host code only.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace fsr {

static const int kTagLen = 30;

static inline void swap_bytes(void* p, size_t m, size_t n)
{
  unsigned char* q = static_cast<unsigned char*>(p);
  for (size_t i = 0; i < n; ++i, q += m)
    for (size_t a = 0, b = m - 1; a < b; ++a, --b) { unsigned char t = q[a]; q[a] = q[b]; q[b] = t; }
}

struct TaggedFile {
  FILE* f = nullptr;
  bool swap = false, ascii = false;
  std::string tag;
  unsigned int checksum = 0;
  ~TaggedFile() { if (f) fclose(f); }

  int open_read(const char* path, bool allow_ascii = false)
  {
    f = fopen(path, "rb");
    if (!f) { set_error("cannot open %s", path); return FSR_ERR_ARG; }
    // FFaTag_read (FFaTag.C:192-250): a line end inside the first 30 characters = an ASCII file (header-only
    // "possibility" files such as FFrTests/response_pos.frs); those have no endian/checksum/version fields
    tag.clear();
    ascii = false;
    bool binary = false;
    for (int i = 0; i < kTagLen; ++i) {
      const int c = fgetc(f);
      if (c == EOF || (i == 0 && c != '#')) { set_error("%s: not a tagged FEDEM binary file", path); return FSR_ERR_ARG; }
      if (!binary && (c == '\n' || c == '\r')) { ascii = true; break; }
      if (c < 32 || c > 126) binary = true;
      tag += (char)c;
    }
    while (!tag.empty() && tag.back() == ' ') tag.pop_back();
    if (ascii) {
      if (!allow_ascii) { set_error("%s: ASCII file where a binary file is expected", path); return FSR_ERR_ARG; }
      return FSR_OK;
    }
    unsigned char e[2];
    unsigned int cs[2];
    if (fread(e, 1, 2, f) != 2 || fread(cs, 4, 2, f) != 2) { set_error("%s: truncated file header", path); return FSR_ERR_ARG; }
    // the writer stores the 16-bit value 0x1234 in its own byte order
    const bool file_little = e[0] == 0x34 && e[1] == 0x12, file_big = e[0] == 0x12 && e[1] == 0x34;
    if (!file_little && !file_big) { set_error("%s: invalid endian field", path); return FSR_ERR_ARG; }
    const uint16_t probe = 0x1234;
    const bool host_little = *reinterpret_cast<const unsigned char*>(&probe) == 0x34;
    swap = file_little != host_little;
    checksum = cs[1];
    if (swap) swap_bytes(&checksum, 4, 1);
    char ver[16];
    if (!fgets(ver, sizeof(ver), f)) { set_error("%s: missing version field", path); return FSR_ERR_ARG; }
    float v = 0.f;
    if (sscanf(ver, ";%f;", &v) < 1 || v != 1.0f) { set_error("%s: wrong file version field '%s'", path, ver); return FSR_ERR_ARG; }
    return FSR_OK;
  }

  int open_write(const char* path, const char* tg, unsigned int cs)
  {
    f = fopen(path, "wb");
    if (!f) { set_error("cannot create %s", path); return FSR_ERR_ARG; }
    char t[kTagLen];
    memset(t, ' ', kTagLen);
    memcpy(t, tg, std::min<size_t>(strlen(tg), (size_t)kTagLen));
    const uint16_t endian = 0x1234;
    const unsigned int c2[2] = {0u, cs};
    if (fwrite(t, 1, kTagLen, f) != (size_t)kTagLen || fwrite(&endian, 2, 1, f) != 1 || fwrite(c2, 4, 2, f) != 2 ||
        fputs(";1.0;\n", f) < 0) { set_error("%s: write error", path); return FSR_ERR_ARG; }
    return FSR_OK;
  }

  template <class T>
  int read(T* p, size_t n, const char* what)
  {
    if (n == 0) return FSR_OK;
    if (fread(p, sizeof(T), n, f) != n) { set_error("unexpected end of file while reading %s", what); return FSR_ERR_ARG; }
    if (swap && sizeof(T) > 1) swap_bytes(p, sizeof(T), n);
    return FSR_OK;
  }
  template <class T>
  int write(const T* p, size_t n)
  {
    if (n && fwrite(p, sizeof(T), n, f) != n) { set_error("write error"); return FSR_ERR_ARG; }
    return FSR_OK;
  }
};

}  // namespace fsr
