// png_writer.h — minimal PNG encoder for the shaded image (8-bit RGBA, zlib "stored" blocks: no compression library
// needed).  The reference shows its frame with glutSwapBuffers and can dump it with glReadPixels
// (MyGLTextureViewer::loadFrameBufferTexture, ShadowMapping/src/Viewers/MyGLTextureViewer.cpp:97-101); this is the
// file-output counterpart of that read-back.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace sgh {

// rgba: H rows of W pixels, row 0 = TOP of the image.  Returns false if the file cannot be written.
bool writePNG(const std::string& path, const uint8_t* rgba, int W, int H);
// float image as the GPU produces it (row 0 = bottom, components in [0,1]) -> 8-bit with the framebuffer's conversion
// (clamp, *255, round to nearest), flipped to top-down
std::vector<uint8_t> toRGBA8TopDown(const float* rgba, int W, int H);

}  // namespace sgh
