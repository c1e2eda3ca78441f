// weighting.h -- reference: icp-variants/weighting.h:8.  applyWeights (weighting.h:39-99) is fused into
// the device matching kernel; only the enum is part of the API surface.
#pragma once
enum { CONSTANT_WEIGHTING = 0, DISTANCES_WEIGHTING, NORMALS_WEIGHTING, COLORS_WEIGHTING };
