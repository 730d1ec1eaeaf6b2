// TEST INFRASTRUCTURE (oracle build only). No-op stand-in for <hdf5.h>: the oracle never writes XDMF/HDF5
// (Source/Output.cpp:75-337 is outside the hot path); these symbols only have to exist for the unity build.
#ifndef WF_SHIM_HDF5_H
#define WF_SHIM_HDF5_H
typedef long long hid_t; typedef unsigned long long hsize_t; typedef int herr_t;
#define H5F_ACC_TRUNC 0
#define H5F_ACC_RDONLY 0
#define H5F_ACC_RD 0
#define H5F_ACC_RDWR 0
#define H5P_DEFAULT 0
#define H5F_SCOPE_GLOBAL 0
inline hid_t  H5Fcreate(const char *, unsigned, hid_t, hid_t) { return 1; }
inline hid_t  H5Fopen(const char *, unsigned, hid_t) { return 1; }
inline herr_t H5Fclose(hid_t) { return 0; }
inline herr_t H5Fflush(hid_t, int) { return 0; }
#endif
