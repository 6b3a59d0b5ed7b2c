// stand-in for <pcl/filters/filter.h> (TEST INFRASTRUCTURE): nothing of it is used by the compiled sources
#include "../point_cloud.h"
