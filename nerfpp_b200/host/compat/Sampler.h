// Forwarding header (src/Sampler.h).
#pragma once
#include "../render_ops.h"
