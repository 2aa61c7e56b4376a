// Forwarding header with the reference's file name (src/LeRF.h): LeRFImpl / LeRF live in the drop-in layer.
#pragma once
#include "../lerf.h"
