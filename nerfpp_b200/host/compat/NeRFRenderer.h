// Forwarding header (src/NeRFRenderer.h).
#pragma once
#include "../renderer.h"
