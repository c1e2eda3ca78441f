// selection.h -- reference: icp-variants/selection.h:8.  The selection itself (all / Bernoulli(p) per
// iteration, selection.h:88-104) runs inside the device loop; only the enum is part of the API surface.
#pragma once
enum { SELECT_ALL = 0, RANDOM_SAMPLING };
