// Forwarding header (src/LibTorchTraining/TorchHeader.h).
#pragma once
#include "../nrf_torch.h"
