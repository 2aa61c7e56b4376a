// Forwarding header (src/RayUtils.h).
#pragma once
#include "../render_ops.h"
