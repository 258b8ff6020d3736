#include "png_writer.h"

#include <cmath>
#include <cstdio>

namespace sgh {

static uint32_t crc_table[256];
static bool crc_ready = false;
static void crcInit() {
  for (uint32_t n = 0; n < 256; n++) {
    uint32_t c = n;
    for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
    crc_table[n] = c;
  }
  crc_ready = true;
}
static uint32_t crc32(const uint8_t* p, size_t n, uint32_t c = 0xFFFFFFFFu) {
  if (!crc_ready) crcInit();
  for (size_t i = 0; i < n; i++) c = crc_table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
  return c;
}
static void put32(std::vector<uint8_t>& v, uint32_t x) { for (int s = 24; s >= 0; s -= 8) v.push_back((uint8_t)(x >> s)); }
static void chunk(std::vector<uint8_t>& out, const char type[4], const std::vector<uint8_t>& data) {
  put32(out, (uint32_t)data.size());
  std::vector<uint8_t> td(type, type + 4);
  td.insert(td.end(), data.begin(), data.end());
  out.insert(out.end(), td.begin(), td.end());
  put32(out, crc32(td.data(), td.size()) ^ 0xFFFFFFFFu);
}

bool writePNG(const std::string& path, const uint8_t* rgba, int W, int H) {
  if (W <= 0 || H <= 0 || !rgba) return false;
  // scanlines with filter type 0
  std::vector<uint8_t> raw;
  raw.reserve((size_t)H * (4 * (size_t)W + 1));
  for (int y = 0; y < H; y++) {
    raw.push_back(0);
    raw.insert(raw.end(), rgba + (size_t)y * W * 4, rgba + (size_t)(y + 1) * W * 4);
  }
  // zlib stream of stored (uncompressed) deflate blocks
  std::vector<uint8_t> z;
  z.push_back(0x78); z.push_back(0x01);
  uint32_t a = 1, b = 0;                               // Adler-32
  size_t pos = 0;
  while (pos < raw.size() || raw.empty()) {
    const size_t n = raw.size() - pos < 65535 ? raw.size() - pos : 65535;
    const bool last = pos + n >= raw.size();
    z.push_back(last ? 1 : 0);
    z.push_back((uint8_t)(n & 0xFF)); z.push_back((uint8_t)(n >> 8));
    z.push_back((uint8_t)(~n & 0xFF)); z.push_back((uint8_t)((~n >> 8) & 0xFF));
    for (size_t i = 0; i < n; i++) { a = (a + raw[pos + i]) % 65521u; b = (b + a) % 65521u; }
    z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
    pos += n;
    if (last) break;
  }
  put32(z, (b << 16) | a);
  std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  std::vector<uint8_t> ihdr;
  put32(ihdr, (uint32_t)W); put32(ihdr, (uint32_t)H);
  ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);   // 8-bit RGBA
  chunk(out, "IHDR", ihdr);
  chunk(out, "IDAT", z);
  chunk(out, "IEND", {});
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) return false;
  const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
  std::fclose(f);
  return ok;
}

std::vector<uint8_t> toRGBA8TopDown(const float* rgba, int W, int H) {
  std::vector<uint8_t> out((size_t)W * H * 4);
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++)
      for (int c = 0; c < 4; c++) {
        float v = rgba[((size_t)(H - 1 - y) * W + x) * 4 + c];
        v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
        if (!(v == v)) v = 0.0f;
        out[((size_t)y * W + x) * 4 + c] = (uint8_t)std::lround(v * 255.0f);
      }
  return out;
}

}  // namespace sgh
