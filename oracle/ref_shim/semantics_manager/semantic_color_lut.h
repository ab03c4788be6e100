// semantics_manager/semantic_color_lut.h — STAND-IN (test infrastructure only).  TopDownMap::Params holds one by value;
// loadSvg (top_down_map.cpp:78-83) asks it for the colour of every class and compares `c[0] << 16 | c[1] << 8 | c[2]` with
// nanosvg's fill colour (0xBBGGRR), so a table entry here is simply that packed value.
#pragma once
#include <array>
#include <cstdint>
#include <vector>
class SemanticColorLut {
 public:
  std::vector<uint32_t> packed;      // per class index
  uint32_t ind2Color(int i) const { return i >= 0 && (size_t)i < packed.size() ? packed[(size_t)i] : 0xFFFFFFFFu; }
  static std::array<uint8_t, 3> unpackColor(uint32_t c) { return {(uint8_t)(c >> 16), (uint8_t)(c >> 8), (uint8_t)c}; }
  template <class A, class B> void color2Ind(const A&, B&) const {}
};
