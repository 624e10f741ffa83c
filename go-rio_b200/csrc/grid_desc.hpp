// The uniform grid of a cloud (plain C++: shared by the CUDA headers and the host-only code).
#pragma once

namespace apd {

struct GridDesc {
  float ox, oy, oz;   // origin (min corner)
  float inv_cell;     // 1 / cell size
  float cell;         // cell size (metres)
  int nx, ny, nz;     // dimensions
};

}  // namespace apd
