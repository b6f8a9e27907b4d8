// jpeg.cpp — JPEG decode for map_Kd textures (stand-in for image 0.25.5 / zune-jpeg 0.4.13).
// Placeholder until the built-in Huffman decoder lands: reports failure so the caller falls
// back to the "<file>.rgba8" sidecar or, like the reference, to the empty texture.
#include "scene.h"

namespace rc {

bool decode_jpeg(const std::vector<uint8_t>&, Image&, std::string* why)
{
    if (why) *why = "JPEG decoder not available (supply a .rgba8 sidecar)";
    return false;
}

}  // namespace rc
