// temporary: entry points not yet implemented
#include "msfl_internal.h"
extern "C" {
int msfl_scan2scan(msfl_engine *, const msfl_cloud *, const msfl_cloud *, const msfl_cloud *, const msfl_cloud *, double *, msfl_stats *) { msfl::set_error("not implemented"); return MSFL_ERR_ARG; }
int msfl_associate_scan(msfl_engine *, const msfl_cloud *, const msfl_cloud *, const msfl_cloud *, const msfl_cloud *, const double *, int32_t *) { msfl::set_error("not implemented"); return MSFL_ERR_ARG; }
int msfl_extract_features(msfl_engine *, const msfl_cloud *, const double *, msfl_features *) { msfl::set_error("not implemented"); return MSFL_ERR_ARG; }
int msfl_voxel_grid(msfl_engine *, const msfl_cloud *, float, float *, size_t *) { msfl::set_error("not implemented"); return MSFL_ERR_ARG; }
}
