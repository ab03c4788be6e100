// semantics_manager/semantic_color_lut.h — STAND-IN (test infrastructure only): TopDownMap::Params holds one by value;
// only the static-map constructor path (not compiled here) uses it.
#pragma once
#include <array>
#include <cstdint>
class SemanticColorLut {
 public:
  uint32_t ind2Color(int) const { return 0; }
  static std::array<uint8_t, 3> unpackColor(uint32_t c) { return {(uint8_t)(c >> 16), (uint8_t)(c >> 8), (uint8_t)c}; }
  template <class A, class B> void color2Ind(const A&, B&) const {}
};
