// Forwarding header (src/CustomOps.h).
#pragma once
#include "../render_ops.h"
