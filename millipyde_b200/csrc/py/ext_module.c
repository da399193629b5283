/*
 * Module `millipyde`: init, module-level functions, shared helpers.
 *
 * Reference: src/millipyde_module.c (functions :70-120, PyInit :133-301).
 * Same names: device_count, get_current_device, best_device, image_from_path,
 * images_from_path, DEVICE_COUNT, and the six types.  Additions are listed in
 * INTEGRATION.md (synchronize, seed, set_semantics, set_fusion, pinned_empty).
 *
 * Import fails with ImportError when no CUDA device can be initialised -- the
 * reference's behaviour (src/millipyde_module.c:146-151) and the "fail loudly"
 * rule: there is no CPU fallback.  MILLIPYDE_NO_DEVICE_OK=1 lets the import
 * succeed on a GPU-less host so that argument validation can be tested there;
 * every call that needs a device then raises RuntimeError.
 */
#define MP_EXT_MAIN
#include "ext_common.h"

static int g_have_devices = 0;

int mpext_have_devices(void) { return g_have_devices; }

int mpext_require_devices(void)
{
    if (g_have_devices) return 0;
    PyErr_SetString(PyExc_RuntimeError, mperr_str(MP_ERROR_NO_DEVICE));
    return -1;
}

PyObject *mpext_raise_status(MPStatus st, const char *where)
{
    const char *detail = mp_last_error();
    if (detail && detail[0])
        PyErr_Format(PyExc_RuntimeError, "%s: %s (%s)", where, mperr_str(st), detail);
    else
        PyErr_Format(PyExc_RuntimeError, "%s: %s", where, mperr_str(st));
    return NULL;
}

static PyObject *mod_device_count(PyObject *self, PyObject *Py_UNUSED(a))
{
    return PyLong_FromLong(g_have_devices ? mpdev_get_device_count() : 0);
}

static PyObject *mod_current_device(PyObject *self, PyObject *Py_UNUSED(a))
{
    int d = mpdev_get_target_device();
    if (d == DEVICE_LOC_NO_AFFINITY) d = mpdev_get_recommended_device();
    return PyLong_FromLong(d);
}

static PyObject *mod_best_device(PyObject *self, PyObject *Py_UNUSED(a))
{
    return PyLong_FromLong(mpdev_get_recommended_device());
}

static PyObject *mod_image_from_path(PyObject *self, PyObject *path) { return mpext_image_from_path(path); }
static PyObject *mod_images_from_path(PyObject *self, PyObject *path) { return mpext_images_from_path(path); }

/* ---- additions ---------------------------------------------------------------- */
static PyObject *mod_synchronize(PyObject *self, PyObject *Py_UNUSED(a))
{
    if (g_have_devices) {
        Py_BEGIN_ALLOW_THREADS
        mpdev_hard_synchronize_all();
        Py_END_ALLOW_THREADS
    }
    Py_RETURN_NONE;
}

static PyObject *mod_seed(PyObject *self, PyObject *arg)
{
    unsigned long long s = PyLong_AsUnsignedLongLong(arg);
    if (PyErr_Occurred()) return NULL;
    mprand_seed(s);
    Py_RETURN_NONE;
}

static PyObject *mod_set_semantics(PyObject *self, PyObject *arg)
{
    const char *s = PyUnicode_AsUTF8(arg);
    if (!s) return NULL;
    if (strcmp(s, "oracle") == 0) mpimg_set_semantics(MP_SEMANTICS_ORACLE);
    else if (strcmp(s, "reference") == 0) mpimg_set_semantics(MP_SEMANTICS_REFERENCE);
    else {
        PyErr_SetString(PyExc_ValueError, "semantics must be 'oracle' or 'reference'");
        return NULL;
    }
    Py_RETURN_NONE;
}

/* fp32 images are defined on [0, 1]; declare 'any' for other float data (include/mp_image.h) */
static PyObject *mod_set_value_range(PyObject *self, PyObject *arg)
{
    const char *s = PyUnicode_AsUTF8(arg);
    if (!s) return NULL;
    if (strcmp(s, "unit") == 0) mpimg_set_value_range(MP_RANGE_UNIT);
    else if (strcmp(s, "any") == 0) mpimg_set_value_range(MP_RANGE_ANY);
    else {
        PyErr_SetString(PyExc_ValueError, "value range must be 'unit' or 'any'");
        return NULL;
    }
    Py_RETURN_NONE;
}

static PyObject *mod_get_semantics(PyObject *self, PyObject *Py_UNUSED(a))
{
    return PyUnicode_FromString(mpimg_get_semantics() == MP_SEMANTICS_REFERENCE ? "reference" : "oracle");
}

static PyObject *mod_set_fusion(PyObject *self, PyObject *arg)
{
    mppipe_set_fusion(PyObject_IsTrue(arg));
    Py_RETURN_NONE;
}

static PyObject *mod_launch_count(PyObject *self, PyObject *Py_UNUSED(a))
{
    return PyLong_FromUnsignedLongLong(mpdev_launch_count());
}

static void pinned_capsule_free(PyObject *capsule)
{
    mphost_free_pinned(PyCapsule_GetPointer(capsule, "mp_pinned"));
}

/* pinned_empty(shape, dtype): ndarray over page-locked host memory, so that gpuimage(arr) and
 * np.array(img) move at full PCIe rate. */
static PyObject *mod_pinned_empty(PyObject *self, PyObject *args)
{
    PyObject *shape_obj, *dtype_obj = NULL;
    if (!PyArg_ParseTuple(args, "O|O", &shape_obj, &dtype_obj)) return NULL;
    if (mpext_require_devices() < 0) return NULL;
    PyArray_Descr *descr = NULL;
    if (!dtype_obj || dtype_obj == Py_None) descr = PyArray_DescrFromType(NPY_FLOAT);
    else if (!PyArray_DescrConverter(dtype_obj, &descr)) return NULL;
    PyArray_Dims dims = {NULL, 0};
    if (!PyArray_IntpConverter(shape_obj, &dims)) {
        Py_DECREF(descr);
        return NULL;
    }
    npy_intp n = PyDataType_ELSIZE(descr);
    for (int i = 0; i < dims.len; ++i) n *= dims.ptr[i];
    void *p = mphost_alloc_pinned((size_t)n);
    if (!p) {
        PyDimMem_FREE(dims.ptr);
        Py_DECREF(descr);
        return PyErr_NoMemory();
    }
    PyObject *arr = PyArray_NewFromDescr(&PyArray_Type, descr, dims.len, dims.ptr, NULL, p,
                                         NPY_ARRAY_CARRAY, NULL); /* steals descr */
    PyDimMem_FREE(dims.ptr);
    if (!arr) {
        mphost_free_pinned(p);
        return NULL;
    }
    PyObject *cap = PyCapsule_New(p, "mp_pinned", pinned_capsule_free);
    if (!cap || PyArray_SetBaseObject((PyArrayObject *)arr, cap) < 0) {
        Py_XDECREF(cap);
        Py_DECREF(arr);
        return NULL;
    }
    return arr;
}

static PyMethodDef module_methods[] = {
    {"device_count", mod_device_count, METH_NOARGS, "number of CUDA devices"},
    {"get_current_device", mod_current_device, METH_NOARGS, "target device, else the recommended one"},
    {"best_device", mod_best_device, METH_NOARGS, "recommended device (max clock x SM count)"},
    {"image_from_path", mod_image_from_path, METH_O, "decode an image file into a gpuimage"},
    {"images_from_path", mod_images_from_path, METH_O, "decode every image of a directory (sorted) into gpuimages"},
    {"synchronize", mod_synchronize, METH_NOARGS, "wait until every device is idle"},
    {"seed", mod_seed, METH_O, "seed the random source of random_* ops and probabilities (0: entropy)"},
    {"set_semantics", mod_set_semantics, METH_O, "'oracle' (scikit-image rules) or 'reference' (kernel-exact)"},
    {"get_semantics", mod_get_semantics, METH_NOARGS, "current semantics"},
    {"set_value_range", mod_set_value_range, METH_O,
     "'unit' (default): float32 images hold [0, 1] data -- gaussian() may use fp16 correction operands and returns "
     "Inf/NaN for |sample| >= 65504; 'any': arbitrary float data, no range limit (FMA-pipe Gaussian)"},
    {"set_fusion", mod_set_fusion, METH_O, "enable/disable the chain fusion pass"},
    {"launch_count", mod_launch_count, METH_NOARGS, "kernels launched by the library so far"},
    {"pinned_empty", mod_pinned_empty, METH_VARARGS, "pinned_empty(shape, dtype=float32): page-locked ndarray"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef millipyde_module = {
    PyModuleDef_HEAD_INIT, "millipyde",
    "GPU image augmentation (B200-native implementation of the Millipyde API)", -1, module_methods};

static void teardown_at_exit(void) { mpdev_teardown(); }

PyMODINIT_FUNC PyInit_millipyde(void)
{
    import_array();

    MPStatus st = mpdev_initialize();
    if (st != MILLIPYDE_SUCCESS) {
        const char *ok = getenv("MILLIPYDE_NO_DEVICE_OK");
        if (!(ok && ok[0] == '1')) {
            PyErr_SetString(PyExc_ImportError, mperr_str(st));
            return NULL;
        }
        g_have_devices = 0;
    } else {
        g_have_devices = 1;
        if (mpdev_get_device_count() > 1 && mpdev_peer_to_peer_supported() == MP_FALSE) {
            if (PyErr_WarnEx(PyExc_ImportWarning, mperr_str(DEV_WARN_NO_PEER_ACCESS), 1) < 0) return NULL;
        }
        Py_AtExit(teardown_at_exit);
    }

    MPImage_Type.tp_base = &MPArray_Type;
    static const struct {
        PyTypeObject *type;
        const char *name;
        MPStatus e_create, e_add;
    } types[] = {
        {&MPArray_Type, "gpuarray", MOD_ERROR_CREATE_GPUARRAY_TYPE, MOD_ERROR_ADD_GPUARRAY},
        {&MPImage_Type, "gpuimage", MOD_ERROR_CREATE_GPUIMAGE_TYPE, MOD_ERROR_ADD_GPUIMAGE},
        {&MPOperation_Type, "Operation", MOD_ERROR_CREATE_OPERATION_TYPE, MOD_ERROR_ADD_OPERATION},
        {&MPPipeline_Type, "Pipeline", MOD_ERROR_CREATE_PIPELINE_TYPE, MOD_ERROR_ADD_PIPELINE},
        {&MPGenerator_Type, "Generator", MOD_ERROR_CREATE_GENERATOR_TYPE, MOD_ERROR_ADD_GENERATOR},
        {&MPDevice_Type, "Device", MOD_ERROR_CREATE_DEVICE_TYPE, MOD_ERROR_ADD_DEVICE},
    };
    for (size_t i = 0; i < sizeof types / sizeof types[0]; ++i) {
        if (PyType_Ready(types[i].type) < 0) {
            PyErr_SetString(PyExc_ImportError, mperr_str(types[i].e_create));
            return NULL;
        }
    }
    PyObject *m = PyModule_Create(&millipyde_module);
    if (!m) return NULL;
    PyModule_AddIntConstant(m, "DEVICE_COUNT", g_have_devices ? mpdev_get_device_count() : 0);
    for (size_t i = 0; i < sizeof types / sizeof types[0]; ++i) {
        Py_INCREF(types[i].type);
        if (PyModule_AddObject(m, types[i].name, (PyObject *)types[i].type) < 0) {
            Py_DECREF(types[i].type);
            Py_DECREF(m);
            PyErr_SetString(PyExc_ImportError, mperr_str(types[i].e_add));
            return NULL;
        }
    }
    return m;
}
