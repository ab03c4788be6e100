// png_gray.hpp — 8-bit grayscale PNG files, the container of the reference's raster cache (`class<i>.png`, written by
// cv::imwrite and read by cv::imread(..., IMREAD_GRAYSCALE), top_down_map.cpp:197-224).  OpenCV is not in this image;
// the format is the PNG specification's, so this is a small codec over zlib (link with -lz): non-interlaced, bit depth
// 8, colour type 0; the writer uses filter 0 on every scan line, the reader undoes all five filter types (OpenCV's
// files use Sub).  The Python twin is top_down_renderer_b200/rastercache.py; tests check both against cv2.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

namespace tdrhost {
namespace png {

inline void put32(std::vector<uint8_t>& v, uint32_t x) { for (int s = 24; s >= 0; s -= 8) v.push_back((uint8_t)(x >> s)); }
inline uint32_t get32(const uint8_t* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }
inline void chunk(std::vector<uint8_t>& out, const char tag[4], const std::vector<uint8_t>& body) {
  put32(out, (uint32_t)body.size());
  const size_t at = out.size();
  out.insert(out.end(), tag, tag + 4);
  out.insert(out.end(), body.begin(), body.end());
  put32(out, (uint32_t)crc32(0L, out.data() + at, (uInt)(out.size() - at)));
}

// img: height x width, row-major, row 0 = top line
inline bool write_gray(const std::string& path, const uint8_t* img, int width, int height) {
  std::vector<uint8_t> raw((size_t)height * (width + 1));
  for (int y = 0; y < height; y++) {
    raw[(size_t)y * (width + 1)] = 0;
    std::copy(img + (size_t)y * width, img + (size_t)(y + 1) * width, raw.begin() + (size_t)y * (width + 1) + 1);
  }
  uLongf bound = compressBound((uLong)raw.size());
  std::vector<uint8_t> z(bound);
  if (compress2(z.data(), &bound, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
  z.resize(bound);
  std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'}, hdr;
  put32(hdr, (uint32_t)width); put32(hdr, (uint32_t)height);
  const uint8_t tail[5] = {8, 0, 0, 0, 0};             // bit depth 8, grayscale, deflate, adaptive filtering, no interlace
  hdr.insert(hdr.end(), tail, tail + 5);
  chunk(out, "IHDR", hdr); chunk(out, "IDAT", z); chunk(out, "IEND", {});
  std::ofstream f(path, std::ios::binary | std::ios::trunc);
  f.write(reinterpret_cast<const char*>(out.data()), (std::streamsize)out.size());
  return (bool)f;
}

inline bool read_gray(const std::string& path, std::vector<uint8_t>& img, int& width, int& height) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  std::vector<uint8_t> d((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
  if (d.size() < 8 || !std::equal(sig, sig + 8, d.begin())) return false;
  std::vector<uint8_t> idat;
  bool have_hdr = false;
  for (size_t pos = 8; pos + 12 <= d.size();) {
    const uint32_t n = get32(&d[pos]);
    if (pos + 12 + (size_t)n > d.size()) return false;
    if (get32(&d[pos + 8 + n]) != (uint32_t)crc32(0L, &d[pos + 4], (uInt)(n + 4))) return false;
    const std::string tag(reinterpret_cast<const char*>(&d[pos + 4]), 4);
    if (tag == "IHDR") {
      if (n != 13) return false;
      width = (int)get32(&d[pos + 8]); height = (int)get32(&d[pos + 12]);
      if (d[pos + 16] != 8 || d[pos + 17] != 0 || d[pos + 20] != 0) return false;   // 8-bit gray, non-interlaced only
      have_hdr = true;
    } else if (tag == "IDAT") {
      idat.insert(idat.end(), d.begin() + pos + 8, d.begin() + pos + 8 + n);
    } else if (tag == "IEND") {
      break;
    }
    pos += 12 + (size_t)n;
  }
  if (!have_hdr || width <= 0 || height <= 0) return false;
  std::vector<uint8_t> raw((size_t)height * (width + 1));
  uLongf len = (uLongf)raw.size();
  if (uncompress(raw.data(), &len, idat.data(), (uLong)idat.size()) != Z_OK || len != raw.size()) return false;
  img.assign((size_t)width * height, 0);
  for (int y = 0; y < height; y++) {
    const uint8_t ft = raw[(size_t)y * (width + 1)];
    const uint8_t* line = &raw[(size_t)y * (width + 1) + 1];
    uint8_t* cur = &img[(size_t)y * width];
    const uint8_t* up = y ? cur - width : nullptr;
    if (ft > 4) return false;
    for (int x = 0; x < width; x++) {
      const int a = x ? cur[x - 1] : 0, b = up ? up[x] : 0, c = (x && up) ? up[x - 1] : 0;
      int pred = 0;
      if (ft == 1) pred = a;
      else if (ft == 2) pred = b;
      else if (ft == 3) pred = (a + b) >> 1;
      else if (ft == 4) {
        const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
        pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
      }
      cur[x] = (uint8_t)(line[x] + pred);
    }
  }
  return true;
}

}  // namespace png
}  // namespace tdrhost
