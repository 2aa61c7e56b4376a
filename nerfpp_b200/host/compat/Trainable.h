// Forwarding header (src/LibTorchTraining/Trainable.h).
#pragma once
#include "../models.h"
