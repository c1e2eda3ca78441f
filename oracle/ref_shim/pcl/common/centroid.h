// oracle/ref_shim: stand-in header (test infrastructure only); everything lives in pcl/point_types.h.
#include <pcl/point_types.h>
