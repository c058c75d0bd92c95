/* TEST INFRASTRUCTURE (oracle/): a stand-in for <hdf5.h>.
 *
 * The image has no libhdf5.  This header declares the subset of the HDF5 C API that the reference engine's
 * sources call (listed in SURVEY.md §2 row 22 / §8(c)); oracle/hdf5_shim/hdf5_shim.cpp implements it on top of
 * the project's own HDF5 reader/writer (upside-md_b200/csrc/h5lite).  With it, the UNMODIFIED reference
 * sources under /root/reference/src compile and run here (oracle/Makefile -> oracle/_ref/).
 * Not part of the product; nothing in the shipped library includes this file.
 */
#ifndef UPSIDE_B200_HDF5_SHIM_H
#define UPSIDE_B200_HDF5_SHIM_H
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t hid_t;
typedef int herr_t;
typedef int htri_t;
typedef unsigned long long hsize_t;
typedef long long hssize_t;

#define H5P_DEFAULT ((hid_t)0)
#define H5E_DEFAULT ((hid_t)0)
#define H5S_ALL ((hid_t)0)
#define H5S_UNLIMITED ((hsize_t)(-1))

#define H5F_ACC_RDONLY 0x0000u
#define H5F_ACC_RDWR 0x0001u
#define H5F_ACC_TRUNC 0x0002u
#define H5F_ACC_EXCL 0x0004u

typedef enum { H5F_SCOPE_LOCAL = 0, H5F_SCOPE_GLOBAL = 1 } H5F_scope_t;
typedef enum { H5S_SCALAR = 0, H5S_SIMPLE = 1, H5S_NULL = 2 } H5S_class_t;
typedef enum { H5S_SELECT_SET = 0 } H5S_seloper_t;
typedef enum { H5T_STR_NULLTERM = 0, H5T_STR_NULLPAD = 1, H5T_STR_SPACEPAD = 2 } H5T_str_t;
typedef enum { H5_INDEX_NAME = 0, H5_INDEX_CRT_ORDER = 1 } H5_index_t;
typedef enum { H5_ITER_INC = 0, H5_ITER_DEC = 1 } H5_iter_order_t;

typedef struct {
    int storage_type;
    hsize_t nlinks;
    int64_t max_corder;
    int mounted;
} H5G_info_t;

/* predefined datatypes are small fixed handle values */
#define H5T_NATIVE_FLOAT ((hid_t)1)
#define H5T_NATIVE_DOUBLE ((hid_t)2)
#define H5T_NATIVE_INT ((hid_t)3)
#define H5T_NATIVE_LONG ((hid_t)4)
#define H5T_NATIVE_UINT ((hid_t)5)
#define H5T_C_S1 ((hid_t)6)
#define H5P_DATASET_CREATE ((hid_t)7)

typedef herr_t (*H5E_auto2_t)(hid_t, void*);

hid_t H5Fopen(const char* path, unsigned flags, hid_t fapl);
hid_t H5Fcreate(const char* path, unsigned flags, hid_t fcpl, hid_t fapl);
herr_t H5Fclose(hid_t f);
herr_t H5Fflush(hid_t obj, H5F_scope_t scope);

hid_t H5Gopen2(hid_t loc, const char* name, hid_t gapl);
hid_t H5Gcreate2(hid_t loc, const char* name, hid_t lcpl, hid_t gcpl, hid_t gapl);
herr_t H5Gclose(hid_t g);
herr_t H5Gget_info_by_name(hid_t loc, const char* name, H5G_info_t* info, hid_t lapl);

htri_t H5Lexists(hid_t loc, const char* name, hid_t lapl);
herr_t H5Ldelete(hid_t loc, const char* name, hid_t lapl);
long H5Lget_name_by_idx(hid_t loc, const char* group, H5_index_t idx, H5_iter_order_t order, hsize_t n, char* name,
                        size_t size, hid_t lapl);
htri_t H5Oexists_by_name(hid_t loc, const char* name, hid_t lapl);

hid_t H5Dopen2(hid_t loc, const char* name, hid_t dapl);
hid_t H5Dcreate2(hid_t loc, const char* name, hid_t dtype, hid_t space, hid_t lcpl, hid_t dcpl, hid_t dapl);
herr_t H5Dclose(hid_t d);
hid_t H5Dget_space(hid_t d);
hid_t H5Dget_type(hid_t d);
herr_t H5Dread(hid_t d, hid_t memtype, hid_t memspace, hid_t filespace, hid_t xfer, void* buf);
herr_t H5Dwrite(hid_t d, hid_t memtype, hid_t memspace, hid_t filespace, hid_t xfer, const void* buf);
herr_t H5Dset_extent(hid_t d, const hsize_t* dims);

hid_t H5Screate(H5S_class_t cls);
hid_t H5Screate_simple(int rank, const hsize_t* dims, const hsize_t* maxdims);
herr_t H5Sclose(hid_t s);
int H5Sget_simple_extent_ndims(hid_t s);
int H5Sget_simple_extent_dims(hid_t s, hsize_t* dims, hsize_t* maxdims);
herr_t H5Sselect_hyperslab(hid_t s, H5S_seloper_t op, const hsize_t* start, const hsize_t* stride,
                           const hsize_t* count, const hsize_t* block);

htri_t H5Aexists_by_name(hid_t loc, const char* obj, const char* attr, hid_t lapl);
hid_t H5Aopen_by_name(hid_t loc, const char* obj, const char* attr, hid_t aapl, hid_t lapl);
hid_t H5Acreate_by_name(hid_t loc, const char* obj, const char* attr, hid_t type, hid_t space, hid_t acpl, hid_t aapl,
                        hid_t lapl);
herr_t H5Aread(hid_t a, hid_t memtype, void* buf);
herr_t H5Awrite(hid_t a, hid_t memtype, const void* buf);
hid_t H5Aget_space(hid_t a);
hid_t H5Aget_type(hid_t a);
herr_t H5Aclose(hid_t a);

hid_t H5Tcopy(hid_t t);
herr_t H5Tset_size(hid_t t, size_t size);
herr_t H5Tset_strpad(hid_t t, H5T_str_t pad);
size_t H5Tget_size(hid_t t);
htri_t H5Tis_variable_str(hid_t t);
herr_t H5Tclose(hid_t t);

hid_t H5Pcreate(hid_t cls);
herr_t H5Pclose(hid_t p);
herr_t H5Pset_chunk(hid_t p, int rank, const hsize_t* dims);
herr_t H5Pset_shuffle(hid_t p);
herr_t H5Pset_fletcher32(hid_t p);
herr_t H5Pset_deflate(hid_t p, unsigned level);

herr_t H5Eset_auto(hid_t stack, H5E_auto2_t func, void* data);
herr_t H5Eprint2(hid_t stack, FILE* stream);
int H5Iinc_ref(hid_t id);

#ifdef __cplusplus
}
#endif
#endif
