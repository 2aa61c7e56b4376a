// Forwarding header with the reference's file name (src/LeRFRenderer.h): LeRFRendererOutputs, LeRFRenderResult, RenderCLIPEmbedding, LeRFRenderer.
#pragma once
#include "../lerf.h"
