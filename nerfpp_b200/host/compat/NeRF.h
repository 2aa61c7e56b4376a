// Forwarding header (src/NeRF.h): Embedder, BaseNeRF, NeRF, NeRFSmall.
#pragma once
#include "../embedders.h"
#include "../models.h"
