// opaque handle behind hsb_format (include/hisparse_b200.h)
#ifndef HISPARSE_B200_FORMAT_HANDLE_H_
#define HISPARSE_B200_FORMAT_HANDLE_H_
#include "tile_format.h"
struct hsb_format {
    hsb::TiledMatrix M;
};
#endif
