// Forwarding header: lets code written against the reference include layout (src/CuHashEmbedder.h) build against nerfpp_b200.
#pragma once
#include "../embedders.h"
