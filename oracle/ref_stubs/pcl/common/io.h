// stand-in for <pcl/common/io.h> (TEST INFRASTRUCTURE): copyPointCloud lives in copy_point.h here
#include "copy_point.h"
