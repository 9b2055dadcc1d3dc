// Constants of the packed BVH layout shared by the host packer (scene_host.cpp) and the kernels.
#pragma once
#include <cstdint>
namespace rayb200 {
// child pointer encodings: inner -> packed index (>= 0); single leaf i -> ~i; leaf pair (i, i+1) -> ~(i | kPairBit)
constexpr int32_t kPairBit = 1 << 30;
constexpr int32_t kLeafIndexMask = kPairBit - 1;
}  // namespace rayb200
