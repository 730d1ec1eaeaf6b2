// TEST INFRASTRUCTURE (oracle build only). No-op stand-in for <hdf5_hl.h>; see hdf5.h.
#ifndef WF_SHIM_HDF5_HL_H
#define WF_SHIM_HDF5_HL_H
#include "hdf5.h"
inline herr_t H5LTmake_dataset_float(hid_t, const char *, int, const hsize_t *, const float *) { return 0; }
inline herr_t H5LTmake_dataset_int(hid_t, const char *, int, const hsize_t *, const int *) { return 0; }
inline herr_t H5LTread_dataset_float(hid_t, const char *, float *) { return 0; }
inline herr_t H5LTread_dataset_int(hid_t, const char *, int *) { return 0; }
#endif
