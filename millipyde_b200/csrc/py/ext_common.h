/*
 * Shared declarations of the `millipyde` CPython extension: the drop-in Python
 * surface (gpuarray, gpuimage, Operation, Pipeline, Generator, Device) over the
 * C ABI of libmp_b200.so.  Mirrors what src/gpuarray.c, gpuimage.c,
 * gpuoperation.c, gpupipeline.c, gpugenerator.c, device.c and
 * millipyde_module.c provide in the reference (same type names, method names,
 * argument conventions and error strings); the arithmetic, scheduling and
 * memory management all live behind the ABI.
 */
#ifndef MP_EXT_COMMON_H
#define MP_EXT_COMMON_H

#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <structmember.h>

#define PY_ARRAY_UNIQUE_SYMBOL mp_b200_ARRAY_API
#ifndef MP_EXT_MAIN
#define NO_IMPORT_ARRAY
#endif
#include <numpy/arrayobject.h>

#include "mp_abi.h"
#include "mp_devices.h"
#include "mp_image.h"
#include "mp_objects.h"
#include "mp_pipeline.h"

/* ---- objects --------------------------------------------------------------- */
typedef struct {
    PyObject_HEAD
    MPObjData *obj; /* owned */
} MPArrayObject;    /* gpuarray and its subclass gpuimage share the layout */

typedef struct {
    PyObject_HEAD
    PyObject *callable;  /* callable or str (instance-method name) */
    PyObject *arg_tuple;
    int requires_instance;
    double probability; /* -1: always */
} MPOperationObject;

typedef struct mp_pipeline_object {
    PyObject_HEAD
    PyObject *inputs;
    PyObject *operations;
    MPPipeline *pipe;
    struct mp_pipeline_object *receiver; /* strong reference */
} MPPipelineObject;

typedef struct {
    PyObject_HEAD
    PyObject *inputs;
    PyObject *operations;
    PyObject *ready;    /* list of produced items not yet handed out */
    MPPipeline *pipe;   /* non-NULL when every op resolves to a C operator */
    int device_id;
    long max;
    long produced;      /* items generated so far (may run ahead of `i`) */
    long i;
    int return_to_host;
    int prefetch;
    /* asynchronous look-ahead (explicit prefetch >= MP_ASYNC_PREFETCH): the batch the devices are working
     * on while the consumer drains `ready`, and what must stay alive until it is done */
    PyObject *inflight;       /* list of the batch's gpuimages, or NULL */
    PyObject *inflight_keep;  /* list of per-batch input replicas the batch's views borrow from */
    Py_ssize_t ready_pos;     /* next item of `ready` to hand out (popping the front of a list is O(n)) */
} MPGeneratorObject;

typedef struct {
    PyObject_HEAD
    int device_id;
    int prev_device_id;
} MPDeviceObject;

extern PyTypeObject MPArray_Type, MPImage_Type, MPOperation_Type, MPPipeline_Type, MPGenerator_Type,
    MPDevice_Type;

#define MP_IS_GPU_OBJECT(o) (PyObject_TypeCheck((PyObject *)(o), &MPArray_Type))

/* ---- helpers shared across the translation units ---------------------------- */
int mpext_have_devices(void); /* 0 when imported with MILLIPYDE_NO_DEVICE_OK=1 on a GPU-less host */
int mpext_require_devices(void); /* sets RuntimeError and returns -1 without devices */
PyObject *mpext_raise_status(MPStatus st, const char *where); /* sets RuntimeError, returns NULL */

PyObject *mpext_to_ndarray(MPArrayObject *self);                   /* D2H into a fresh ndarray */
PyObject *mpext_wrap_obj(PyTypeObject *type, MPObjData *obj);      /* takes ownership of obj */
PyObject *mpext_clone(MPArrayObject *self, int device_id, int stream_id);
PyObject *mpext_image_from_path(PyObject *path);
PyObject *mpext_images_from_path(PyObject *path);

/* Resolve an Operation to a C stage: returns 1 and fills *out (args malloc'ed, caller frees),
 * 0 if the Operation is not a string-named C operator, -1 with an exception set on bad args. */
int mpext_resolve_operation(MPOperationObject *op, MPRunnable *out);
PyObject *mpext_operation_run_on(MPOperationObject *op, PyObject *instance);
PyObject *mpext_operation_run(MPOperationObject *op);

#endif /* MP_EXT_COMMON_H */
