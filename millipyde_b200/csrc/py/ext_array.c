/*
 * gpuarray and gpuimage.
 *
 * Reference: src/gpuarray.c (ctor :41-117, __array__ :120-144, numpy protocols
 * :147-226, clone :244-275) and src/gpuimage.c (ctor :18-73, the eight ops and
 * their random_* variants :88-514, clone :516-565, loaders :575-688, argument
 * parsers :691-781).  Same names, argument formats and error strings.  Ops
 * mutate in place and return None; unlike the reference, an MPStatus other than
 * success raises RuntimeError instead of being dropped.
 */
#include <math.h>
#include <strings.h>
#include "ext_common.h"

/* ------------------------------------------------------------------ lifecycle */
static void array_dealloc(MPArrayObject *self)
{
    if (self->obj) {
        if (mpext_have_devices()) mpobj_dealloc_device_data(self->obj);
        free(self->obj->dims);
        free(self->obj);
    }
    Py_TYPE(self)->tp_free((PyObject *)self);
}

static PyObject *array_new(PyTypeObject *type, PyObject *args, PyObject *kwds)
{
    MPArrayObject *self = (MPArrayObject *)type->tp_alloc(type, 0);
    if (self) self->obj = NULL;
    return (PyObject *)self;
}

/* Shared by gpuarray.__init__ and gpuimage.__init__; `image` selects the error
 * strings and the 2-D/3-D rule (src/gpuimage.c:49-63). */
static int init_from_any(MPArrayObject *self, PyObject *args, int image)
{
    PyObject *any = NULL;
    if (!PyArg_ParseTuple(args, "O", &any)) return -1;
    const MPStatus e_type = image ? GPUIMAGE_ERROR_CONSTRUCTION_WITHOUT_ARRAY_TYPE
                                  : GPUARRAY_ERROR_CONSTRUCTION_WITHOUT_ARRAY_TYPE;
    const MPStatus e_num = image ? GPUIMAGE_ERROR_CONSTRUCTION_WITHOUT_IMAGE_FORMAT
                                 : GPUARRAY_ERROR_CONSTRUCTION_WITHOUT_NUMERIC_ARRAY;
    PyArrayObject *array =
        (PyArrayObject *)PyArray_FROM_OTF(any, NPY_NOTYPE, NPY_ARRAY_IN_ARRAY); /* new ref, C-contiguous */
    if (array == NULL) {
        PyErr_Clear();
        PyErr_SetString(PyExc_ValueError, mperr_str(e_type));
        return -1;
    }
    if (!PyArray_ISNUMBER(array)) {
        Py_DECREF(array);
        PyErr_SetString(PyExc_ValueError, mperr_str(e_num));
        return -1;
    }
    const int ndim = PyArray_NDIM(array);
    if (image && ndim != 2 && ndim != 3) {
        Py_DECREF(array);
        PyErr_SetString(PyExc_ValueError, mperr_str(GPUIMAGE_ERROR_CONSTRUCTION_WITHOUT_IMAGE_DIMS));
        return -1;
    }
    if (mpext_require_devices() < 0) {
        Py_DECREF(array);
        return -1;
    }
    if (self->obj) { /* re-init */
        mpobj_dealloc_device_data(self->obj);
        free(self->obj->dims);
        free(self->obj);
        self->obj = NULL;
    }
    MPObjData *o = (MPObjData *)calloc(1, sizeof(MPObjData));
    o->mem_loc = HOST_LOC;
    o->type = -1;
    const size_t nbytes = (size_t)PyArray_NBYTES(array);
    mpobj_copy_from_host(o, PyArray_DATA(array), nbytes); /* H2D on the target / recommended device */
    if (nbytes && !o->device_data) {
        free(o);
        Py_DECREF(array);
        mpext_raise_status(MP_ERROR_DEVICE_ALLOC, "gpuarray()");
        return -1;
    }
    o->stream = mpdev_get_stream(o->mem_loc, 0);
    o->ndims = ndim;
    o->type = PyArray_TYPE(array);
    const int slots = ndim < 3 ? 3 : ndim; /* rgb2grey shrinks 3 -> 2; nothing ever grows it */
    o->dims = (int *)calloc(2 * (size_t)slots, sizeof(int));
    for (int i = 0; i < ndim; ++i) {
        o->dims[i] = (int)PyArray_DIMS(array)[i];
        o->dims[i + ndim] = (int)PyArray_STRIDES(array)[i];
    }
    self->obj = o;
    Py_DECREF(array);
    return 0;
}

static int array_init(MPArrayObject *self, PyObject *args, PyObject *kwds) { return init_from_any(self, args, 0); }
static int image_init(MPArrayObject *self, PyObject *args, PyObject *kwds) { return init_from_any(self, args, 1); }

PyObject *mpext_wrap_obj(PyTypeObject *type, MPObjData *obj)
{
    if (!obj) return mpext_raise_status(MP_ERROR_DEVICE_ALLOC, "clone");
    MPArrayObject *r = (MPArrayObject *)type->tp_alloc(type, 0);
    if (!r) {
        mpobj_destroy(obj);
        return NULL;
    }
    r->obj = obj;
    return (PyObject *)r;
}

/* --------------------------------------------------------------------- to host */
PyObject *mpext_to_ndarray(MPArrayObject *self)
{
    MPObjData *o = self->obj;
    if (!o) {
        PyErr_SetString(PyExc_ValueError, "uninitialised gpuarray");
        return NULL;
    }
    npy_intp dims[NPY_MAXDIMS];
    for (int i = 0; i < o->ndims; ++i) dims[i] = o->dims[i];
    PyObject *arr = PyArray_SimpleNew(o->ndims, dims, o->type); /* owns its buffer */
    if (!arr) return NULL;
    MPStatus st = MILLIPYDE_SUCCESS;
    if (o->nbytes) {
        void *dst = PyArray_DATA((PyArrayObject *)arr);
        Py_BEGIN_ALLOW_THREADS
        st = mpobj_copy_to_host_into(o, dst, o->nbytes);
        Py_END_ALLOW_THREADS
    }
    if (st != MILLIPYDE_SUCCESS) {
        Py_DECREF(arr);
        return mpext_raise_status(st, "__array__");
    }
    return arr;
}

/* __array__(dtype=None, copy=None): numpy >= 2 passes both. */
static PyObject *array_array(MPArrayObject *self, PyObject *args, PyObject *kwds)
{
    PyObject *dtype = Py_None, *copy = Py_None;
    static char *kw[] = {"dtype", "copy", NULL};
    if (!PyArg_ParseTupleAndKeywords(args, kwds, "|OO", kw, &dtype, &copy)) return NULL;
    PyObject *arr = mpext_to_ndarray(self);
    if (!arr || dtype == Py_None) return arr;
    PyObject *cast = PyObject_CallMethod(arr, "astype", "O", dtype);
    Py_DECREF(arr);
    return cast;
}

/* Replace every gpuarray/gpuimage in a sequence by its host copy. */
static PyObject *hostify_tuple(PyObject *seq)
{
    Py_ssize_t n = PySequence_Size(seq);
    if (n < 0) return NULL;
    PyObject *out = PyTuple_New(n);
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject *item = PySequence_GetItem(seq, i); /* new ref */
        if (!item) {
            Py_DECREF(out);
            return NULL;
        }
        if (MP_IS_GPU_OBJECT(item)) {
            PyObject *h = mpext_to_ndarray((MPArrayObject *)item);
            Py_DECREF(item);
            if (!h) {
                Py_DECREF(out);
                return NULL;
            }
            item = h;
        }
        PyTuple_SET_ITEM(out, i, item);
    }
    return out;
}

static int scalar_as_double(PyObject *o, double *out);

/* numpy functions that ARE one of the device operators: np.fliplr(img) and
 * np.transpose(img) [2-D, or HWC with axes=(1, 0, 2)] run the CUDA kernel on a device-side copy and
 * return a new object of the caller's type -- no D2H, no CPU work; the source is not modified (numpy
 * functions do not mutate).  Returns a new reference, or NULL with *handled = 0 when the call is not
 * one of these (other function, other axes, layout the operator does not take). */
static PyObject *device_function(MPArrayObject *self, PyObject *func, PyObject *fargs, PyObject *fkw, int *handled)
{
    *handled = 0;
    if (!self->obj || !self->obj->device_data || !PyTuple_Check(fargs) || PyTuple_GET_SIZE(fargs) < 1 ||
        PyTuple_GET_ITEM(fargs, 0) != (PyObject *)self)
        return NULL;
    PyObject *name = PyObject_GetAttrString(func, "__name__");
    if (!name) {
        PyErr_Clear();
        return NULL;
    }
    const char *fn = PyUnicode_Check(name) ? PyUnicode_AsUTF8(name) : NULL;
    MPFunc op = NULL;
    const Py_ssize_t nargs = PyTuple_GET_SIZE(fargs);
    const int has_kw = fkw && fkw != Py_None && PyDict_Check(fkw) && PyDict_Size(fkw) > 0;
    if (fn && strcmp(fn, "fliplr") == 0 && nargs == 1 && !has_kw && self->obj->ndims >= 2) {
        op = mpimg_fliplr;
    } else if (fn && strcmp(fn, "transpose") == 0 && nargs <= 2) {
        PyObject *axes = nargs == 2 ? PyTuple_GET_ITEM(fargs, 1) : NULL;
        if (has_kw) {
            PyObject *k = PyDict_GetItemString(fkw, "axes");
            if (k && PyDict_Size(fkw) == 1 && !axes) axes = k;
            else axes = Py_Ellipsis; /* anything else: not ours */
        }
        const int nd = self->obj->ndims;
        if (!axes || axes == Py_None) {
            if (nd == 2) op = mpimg_transpose;   /* numpy reverses ALL axes: ours only for 2-D */
        } else if (PyTuple_Check(axes) && PyTuple_GET_SIZE(axes) == nd && (nd == 2 || nd == 3)) {
            long want[3] = {1, 0, 2};
            int same = 1;
            for (int i = 0; i < nd; ++i) {
                PyObject *a = PyTuple_GET_ITEM(axes, i);
                if (!PyLong_Check(a) || PyLong_AsLong(a) != want[i]) same = 0;
            }
            if (same) op = mpimg_transpose;
        }
    }
    /* np.clip(img, lo, hi) with scalar bounds on a float32 image: numpy routes the *function* np.clip
     * here (not the clip ufunc), so it is recognised at this level and run as the elementwise kernel */
    ElementwiseArgs ew = {0, 0, 0, 0, 0};
    void *op_args = NULL;
    if (!op && fn && strcmp(fn, "clip") == 0 && nargs == 3 && !has_kw && self->obj->type == NPY_FLOAT) {
        double lo, hi;
        if (scalar_as_double(PyTuple_GET_ITEM(fargs, 1), &lo) && scalar_as_double(PyTuple_GET_ITEM(fargs, 2), &hi)) {
            ew.kind = MP_EW_CLIP;
            ew.a = lo;
            ew.b = hi;
            op = mpimg_elementwise;
            op_args = &ew;
        }
    }
    Py_DECREF(name);
    if (!op) return NULL;
    MPArrayObject *copy = (MPArrayObject *)mpext_clone(self, self->obj->mem_loc, 0);
    if (!copy) {
        PyErr_Clear();
        return NULL;
    }
    MPStatus st;
    Py_BEGIN_ALLOW_THREADS
    st = op(copy->obj, op_args);
    Py_END_ALLOW_THREADS
    if (st != MILLIPYDE_SUCCESS) { /* e.g. an integer gpuarray: the host path below handles it */
        Py_DECREF(copy);
        return NULL;
    }
    *handled = 1;
    return (PyObject *)copy;
}

/* __array_function__(func, types, args, kwargs): device operators where numpy's function is one
 * (above), everything else on host copies as src/gpuarray.c:194-226. */
static PyObject *array_function(MPArrayObject *self, PyObject *args, PyObject *kwds)
{
    PyObject *func, *types, *fargs, *fkw;
    if (!PyArg_ParseTuple(args, "OOOO", &func, &types, &fargs, &fkw)) return NULL;
    int handled = 0;
    PyObject *on_device = device_function(self, func, fargs, fkw, &handled);
    if (handled) return on_device;
    PyObject *host_args = hostify_tuple(fargs);
    if (!host_args) return NULL;
    PyObject *res = PyObject_Call(func, host_args, (fkw == Py_None) ? NULL : fkw);
    Py_DECREF(host_args);
    return res;
}

/* A Python / numpy scalar as a double; 0 if `o` is not a real scalar (arrays, gpuarrays, complex). */
static int scalar_as_double(PyObject *o, double *out)
{
    if (PyBool_Check(o)) return 0;   /* numpy would promote differently: leave it to the host path */
    /* Python floats / ints are "weak" scalars (NEP 50): the float32 image keeps float32.  numpy scalars
     * carry their type: only float32 / float16 leave the result float32 (np.float64(0.3) promotes). */
    if ((PyFloat_Check(o) && !PyArray_IsScalar(o, Double)) || PyLong_Check(o) || PyArray_IsScalar(o, Float) ||
        PyArray_IsScalar(o, Half)) {
        double v = PyFloat_AsDouble(o);
        if (v == -1.0 && PyErr_Occurred()) {
            PyErr_Clear();
            return 0;
        }
        *out = v;
        return 1;
    }
    return 0;
}

/* A per-channel factor array: a 1-D float32 ndarray of three values (a list or a float64 array would
 * promote the float32 image to float64 in numpy: host path). */
static int triple_as_doubles(PyObject *o, double *out)
{
    if (MP_IS_GPU_OBJECT(o) || !PyArray_Check(o) || PyArray_TYPE((PyArrayObject *)o) != NPY_FLOAT) return 0;
    PyArrayObject *a = (PyArrayObject *)PyArray_FROM_OTF(o, NPY_DOUBLE, NPY_ARRAY_IN_ARRAY);
    if (!a) {
        PyErr_Clear();
        return 0;
    }
    int ok = PyArray_NDIM(a) == 1 && PyArray_DIM(a, 0) == 3;
    for (int i = 0; ok && i < 3; ++i) out[i] = ((double *)PyArray_DATA(a))[i];
    Py_DECREF(a);
    return ok;
}

/* ufuncs that ARE a pointwise device kernel on a float32 image: np.add, np.subtract,
 * np.multiply (scalar, or three per-channel factors on an RGB image), np.power (scalar exponent),
 * np.clip, np.maximum, np.minimum with scalar bounds.  The result is a NEW object of the caller's
 * type on the same device -- no D2H, no CPU arithmetic; the source is not modified.  Everything else
 * (other ufuncs, methods other than __call__, out=, dtype=, where=, array operands, non-float32
 * layouts) returns NULL with *handled = 0 and takes the host path like src/gpuarray.c:147-191. */
static PyObject *device_ufunc(MPArrayObject *self, PyObject *ufunc, PyObject *method, PyObject *inputs, PyObject *kwds,
                              int *handled)
{
    *handled = 0;
    if (!self->obj || !self->obj->device_data || self->obj->type != NPY_FLOAT) return NULL;
    if (kwds && PyDict_Size(kwds) > 0) {
        /* np.clip(...) always passes out=None along; any real keyword (out array, dtype, where, casting) is the host's */
        PyObject *out = PyDict_GetItemString(kwds, "out");
        if (PyDict_Size(kwds) != 1 || !out) return NULL;
        if (out != Py_None && !(PyTuple_Check(out) && PyTuple_GET_SIZE(out) == 1 && PyTuple_GET_ITEM(out, 0) == Py_None))
            return NULL;
    }
    if (!PyUnicode_Check(method) || strcmp(PyUnicode_AsUTF8(method), "__call__") != 0) return NULL;
    PyObject *name = PyObject_GetAttrString(ufunc, "__name__");
    if (!name) {
        PyErr_Clear();
        return NULL;
    }
    const char *fn = PyUnicode_Check(name) ? PyUnicode_AsUTF8(name) : "";
    const Py_ssize_t n = PyTuple_GET_SIZE(inputs);
    PyObject *x0 = n > 0 ? PyTuple_GET_ITEM(inputs, 0) : NULL, *x1 = n > 1 ? PyTuple_GET_ITEM(inputs, 1) : NULL;
    ElementwiseArgs ew = {0, 0, 0, 0, 0};
    double v = 0, t[3];
    int ok = 0;
    const int self_first = x0 == (PyObject *)self, self_second = x1 == (PyObject *)self;
    const int channels = self->obj->ndims == 3 ? self->obj->dims[2] : 1;
    if (n == 2 && (strcmp(fn, "add") == 0 || strcmp(fn, "multiply") == 0) && (self_first || self_second)) {
        PyObject *other = self_first ? x1 : x0;          /* commutative */
        const int mul = fn[0] == 'm';
        if (scalar_as_double(other, &v)) {
            ew.kind = mul ? MP_EW_MUL : MP_EW_ADD;
            ew.a = v;
            ok = 1;
        } else if (mul && channels == 3 && triple_as_doubles(other, t)) {
            ew.kind = MP_EW_MUL;
            ew.a = t[0]; ew.b = t[1]; ew.c = t[2];
            ew.per_channel = 1;
            ok = 1;
        }
    } else if (n == 2 && strcmp(fn, "subtract") == 0 && self_first && scalar_as_double(x1, &v)) {
        ew.kind = MP_EW_ADD;      /* x - s == x + (-s) exactly in IEEE arithmetic */
        ew.a = -v;
        ok = 1;
    } else if (n == 2 && strcmp(fn, "power") == 0 && self_first && scalar_as_double(x1, &v)) {
        ew.kind = MP_EW_POW;
        ew.a = v;
        ok = 1;
    } else if (n == 3 && strcmp(fn, "clip") == 0 && self_first && scalar_as_double(x1, &v) &&
               scalar_as_double(PyTuple_GET_ITEM(inputs, 2), &t[0])) {
        ew.kind = MP_EW_CLIP;
        ew.a = v; ew.b = t[0];
        ok = 1;
    } else if (n == 2 && (strcmp(fn, "maximum") == 0 || strcmp(fn, "minimum") == 0) && (self_first || self_second) &&
               scalar_as_double(self_first ? x1 : x0, &v) && v == v) {
        ew.kind = MP_EW_CLIP;     /* one-sided clip (NaN bounds propagate in numpy: host path) */
        ew.a = fn[1] == 'a' ? v : -INFINITY;
        ew.b = fn[1] == 'a' ? INFINITY : v;
        ok = 1;
    }
    Py_DECREF(name);
    /* the scalar must be exactly representable in the image's float32 arithmetic, as numpy's weak
     * scalar promotion makes it (a Python float operand is cast to float32 before the loop runs) */
    if (!ok) return NULL;
    MPArrayObject *copy = (MPArrayObject *)mpext_clone(self, self->obj->mem_loc, 0);
    if (!copy) {
        PyErr_Clear();
        return NULL;
    }
    MPStatus st;
    Py_BEGIN_ALLOW_THREADS
    st = mpimg_elementwise(copy->obj, &ew);
    Py_END_ALLOW_THREADS
    if (st != MILLIPYDE_SUCCESS) {
        Py_DECREF(copy);
        return NULL;
    }
    *handled = 1;
    return (PyObject *)copy;
}

/* __array_ufunc__(ufunc, method, *inputs, **kwargs): the ufuncs above run on the device; the rest
 * follows the reference's host-copy rule (its own version is a printing stub, src/gpuarray.c:147-191). */
static PyObject *array_ufunc(MPArrayObject *self, PyObject *args, PyObject *kwds)
{
    if (PyTuple_Size(args) < 2) {
        PyErr_SetString(PyExc_TypeError, "__array_ufunc__ needs (ufunc, method, *inputs)");
        return NULL;
    }
    PyObject *ufunc = PyTuple_GetItem(args, 0), *method = PyTuple_GetItem(args, 1);
    PyObject *inputs = PyTuple_GetSlice(args, 2, PyTuple_Size(args));
    {
        int handled = 0;
        PyObject *on_device = device_ufunc(self, ufunc, method, inputs, kwds, &handled);
        if (handled) {
            Py_DECREF(inputs);
            return on_device;
        }
        if (PyErr_Occurred()) PyErr_Clear();
    }
    PyObject *host_in = hostify_tuple(inputs);
    Py_DECREF(inputs);
    if (!host_in) return NULL;
    PyObject *bound = PyObject_GetAttr(ufunc, method);
    if (!bound) {
        Py_DECREF(host_in);
        return NULL;
    }
    if (kwds && PyDict_GetItemString(kwds, "out")) { /* out=(gpuarray,) cannot alias host memory */
        Py_DECREF(bound);
        Py_DECREF(host_in);
        Py_RETURN_NOTIMPLEMENTED;
    }
    PyObject *res = PyObject_Call(bound, host_in, kwds);
    Py_DECREF(bound);
    Py_DECREF(host_in);
    return res;
}

/* ------------------------------------------------------------------------ clone */
PyObject *mpext_clone(MPArrayObject *self, int device_id, int stream_id)
{
    if (!self->obj) {
        PyErr_SetString(PyExc_ValueError, "uninitialised gpuarray");
        return NULL;
    }
    MPObjData *c;
    Py_BEGIN_ALLOW_THREADS
    c = mpobj_clone_data(self->obj, device_id, stream_id);
    Py_END_ALLOW_THREADS
    return mpext_wrap_obj(Py_TYPE(self), c);
}

/* Deep copy on the target device, else the recommended one (src/gpuarray.c:244-254). */
static PyObject *array_clone(MPArrayObject *self, PyObject *Py_UNUSED(ignored))
{
    int dev = mpdev_get_target_device();
    if (dev == DEVICE_LOC_NO_AFFINITY) dev = mpdev_get_recommended_device();
    return mpext_clone(self, dev, 0);
}

/* ------------------------------------------------------------------- image ops */
/* Every op: hop to the target device first if one is set and the image is not
 * pinned (src/gpuimage.c:93-103, repeated in each method there). */
static void follow_target_device(MPObjData *o)
{
    if (o->pinned) return;
    const int target = mpdev_get_target_device();
    if (target != DEVICE_LOC_NO_AFFINITY && target != o->mem_loc) mpobj_change_device(o, target);
}

static PyObject *run_op(MPArrayObject *self, MPFunc fn, void *op_args, const char *name)
{
    if (!self->obj) {
        PyErr_SetString(PyExc_ValueError, "uninitialised gpuimage");
        return NULL;
    }
    follow_target_device(self->obj);
    MPStatus st = fn(self->obj, op_args);
    if (st != MILLIPYDE_SUCCESS) return mpext_raise_status(st, name);
    Py_RETURN_NONE;
}

static PyObject *image_grey(MPArrayObject *self, PyObject *Py_UNUSED(ignored))
{
    return run_op(self, mpimg_color_to_greyscale, NULL, "rgb2grey");
}
static PyObject *image_transpose(MPArrayObject *self, PyObject *Py_UNUSED(ignored))
{
    return run_op(self, mpimg_transpose, NULL, "transpose");
}
static PyObject *image_fliplr(MPArrayObject *self, PyObject *Py_UNUSED(ignored))
{
    return run_op(self, mpimg_fliplr, NULL, "fliplr");
}
static PyObject *image_rotate(MPArrayObject *self, PyObject *args)
{
    RotateArgs a;
    if (!PyArg_ParseTuple(args, "d", &a.angle)) return NULL;
    return run_op(self, mpimg_rotate, &a, "rotate");
}
static PyObject *image_gaussian(MPArrayObject *self, PyObject *args)
{
    GaussianArgs a;
    if (!PyArg_ParseTuple(args, "d", &a.sigma)) return NULL;
    return run_op(self, mpimg_gaussian, &a, "gaussian");
}
static PyObject *image_brightness(MPArrayObject *self, PyObject *args)
{
    BrightnessArgs a;
    if (!PyArg_ParseTuple(args, "d", &a.delta)) return NULL;
    if (a.delta <= -1 || a.delta >= 1) { /* src/gpuimage.c:729 (there: NULL without an exception) */
        PyErr_SetString(PyExc_ValueError, "brightness delta must lie strictly between -1 and 1");
        return NULL;
    }
    return run_op(self, mpimg_brightness, &a, "brightness");
}
static PyObject *image_gamma(MPArrayObject *self, PyObject *args)
{
    GammaArgs a;
    if (!PyArg_ParseTuple(args, "dd", &a.gamma, &a.gain)) return NULL;
    return run_op(self, mpimg_adjust_gamma, &a, "adjust_gamma");
}
static PyObject *image_colorize(MPArrayObject *self, PyObject *args)
{
    ColorizeArgs a;
    if (!PyArg_ParseTuple(args, "ddd", &a.r_mult, &a.g_mult, &a.b_mult)) return NULL;
    if (a.r_mult < 0 || a.g_mult < 0 || a.b_mult < 0) { /* src/gpuimage.c:752 */
        PyErr_SetString(PyExc_ValueError, "colorize multipliers must be >= 0");
        return NULL;
    }
    return run_op(self, mpimg_colorize, &a, "colorize");
}

/* random_*: same argument formats as src/gpuimage.c:206, :272, :335, :398, :475. */
static PyObject *image_rand_range(MPArrayObject *self, PyObject *args, MPFunc fn, const char *name)
{
    RandomRangeArgs r;
    if (!PyArg_ParseTuple(args, "dd", &r.min, &r.max)) return NULL;
    return run_op(self, fn, &r, name);
}
static PyObject *image_rand_rotate(MPArrayObject *self, PyObject *args)
{
    return image_rand_range(self, args, mpimg_random_rotate, "random_rotate");
}
static PyObject *image_rand_gaussian(MPArrayObject *self, PyObject *args)
{
    return image_rand_range(self, args, mpimg_random_gaussian, "random_gaussian");
}
static PyObject *image_rand_brightness(MPArrayObject *self, PyObject *args)
{
    return image_rand_range(self, args, mpimg_random_brightness, "random_brightness");
}

static int pair_from_list(PyObject *list, double *lo, double *hi)
{
    if (PyList_Size(list) != 2) {
        PyErr_SetString(PyExc_ValueError, "expected a [min, max] list");
        return -1;
    }
    *lo = PyFloat_AsDouble(PyList_GetItem(list, 0));
    *hi = PyFloat_AsDouble(PyList_GetItem(list, 1));
    return PyErr_Occurred() ? -1 : 0;
}

static PyObject *image_rand_gamma(MPArrayObject *self, PyObject *args)
{
    PyObject *g, *k;
    RandomGammaArgs r;
    if (!PyArg_ParseTuple(args, "O!O!", &PyList_Type, &g, &PyList_Type, &k)) return NULL;
    if (pair_from_list(g, &r.gamma_min, &r.gamma_max) < 0 || pair_from_list(k, &r.gain_min, &r.gain_max) < 0)
        return NULL;
    return run_op(self, mpimg_random_adjust_gamma, &r, "random_adjust_gamma");
}

static PyObject *image_rand_colorize(MPArrayObject *self, PyObject *args)
{
    PyObject *rr, *gg, *bb;
    RandomColorizeArgs r;
    if (!PyArg_ParseTuple(args, "O!O!O!", &PyList_Type, &rr, &PyList_Type, &gg, &PyList_Type, &bb)) return NULL;
    if (pair_from_list(rr, &r.r_min, &r.r_max) < 0 || pair_from_list(gg, &r.g_min, &r.g_max) < 0 ||
        pair_from_list(bb, &r.b_min, &r.b_max) < 0)
        return NULL;
    return run_op(self, mpimg_random_colorize, &r, "random_colorize");
}

/* ---------------------------------------------------------------------- loaders */
/* The reference calls skimage.io.imread through the C API (src/gpuimage.c:575-640);
 * scikit-image is not a dependency here: PIL decodes, numpy wraps. */
PyObject *mpext_image_from_path(PyObject *path)
{
    PyObject *pil = PyImport_ImportModule("PIL.Image");
    if (!pil) return NULL;
    PyObject *img = PyObject_CallMethod(pil, "open", "O", path);
    Py_DECREF(pil);
    if (!img) return NULL;
    PyObject *np = PyImport_ImportModule("numpy");
    PyObject *arr = np ? PyObject_CallMethod(np, "asarray", "O", img) : NULL;
    Py_XDECREF(np);
    Py_DECREF(img);
    if (!arr) return NULL;
    PyObject *res = PyObject_CallFunctionObjArgs((PyObject *)&MPImage_Type, arr, NULL);
    Py_DECREF(arr);
    return res;
}

static int valid_image_filename(const char *name)
{
    const char *ext = strrchr(name, '.');
    if (!ext || ext == name) return 0;
    static const char *ok[] = {"png", "jpg", "jpeg", "tiff", "bmp"}; /* src/gpuimage.c:803-823 */
    for (size_t i = 0; i < sizeof ok / sizeof ok[0]; ++i)
        if (strcasecmp(ext + 1, ok[i]) == 0) return 1;
    return 0;
}

/* Every image file of a directory, in sorted name order (the reference uses readdir order,
 * which is filesystem-dependent). */
PyObject *mpext_images_from_path(PyObject *path)
{
    PyObject *os = PyImport_ImportModule("os");
    if (!os) return NULL;
    PyObject *names = PyObject_CallMethod(os, "listdir", "O", path);
    if (!names) {
        Py_DECREF(os);
        return NULL;
    }
    if (PyList_Sort(names) < 0) {
        Py_DECREF(names);
        Py_DECREF(os);
        return NULL;
    }
    PyObject *ospath = PyObject_GetAttrString(os, "path");
    PyObject *result = PyList_New(0);
    for (Py_ssize_t i = 0; ospath && i < PyList_Size(names); ++i) {
        PyObject *name = PyList_GetItem(names, i);
        const char *cname = PyUnicode_AsUTF8(name);
        if (!cname || !valid_image_filename(cname)) {
            PyErr_Clear();
            continue;
        }
        PyObject *full = PyObject_CallMethod(ospath, "join", "OO", path, name);
        PyObject *isfile = full ? PyObject_CallMethod(ospath, "isfile", "O", full) : NULL;
        if (isfile && PyObject_IsTrue(isfile)) {
            PyObject *img = mpext_image_from_path(full);
            if (!img) {
                Py_XDECREF(isfile);
                Py_XDECREF(full);
                Py_CLEAR(result);
                break;
            }
            PyList_Append(result, img);
            Py_DECREF(img);
        }
        Py_XDECREF(isfile);
        Py_XDECREF(full);
    }
    Py_XDECREF(ospath);
    Py_DECREF(names);
    Py_DECREF(os);
    return result;
}

/* ------------------------------------------------------------------ attributes */
static PyObject *array_get_shape(MPArrayObject *self, void *closure)
{
    if (!self->obj) Py_RETURN_NONE;
    PyObject *t = PyTuple_New(self->obj->ndims);
    for (int i = 0; i < self->obj->ndims; ++i) PyTuple_SET_ITEM(t, i, PyLong_FromLong(self->obj->dims[i]));
    return t;
}
static PyObject *array_get_device(MPArrayObject *self, void *closure)
{
    return PyLong_FromLong(self->obj ? self->obj->mem_loc : HOST_LOC);
}
static PyObject *array_get_dtype(MPArrayObject *self, void *closure)
{
    if (!self->obj) Py_RETURN_NONE;
    return (PyObject *)PyArray_DescrFromType(self->obj->type);
}
static PyObject *image_get_width(MPArrayObject *self, void *closure)
{
    return PyLong_FromLong(self->obj && self->obj->ndims > 1 ? self->obj->dims[1] : 0);
}
static PyObject *image_get_height(MPArrayObject *self, void *closure)
{
    return PyLong_FromLong(self->obj && self->obj->ndims > 0 ? self->obj->dims[0] : 0);
}

static PyGetSetDef array_getset[] = {
    {"shape", (getter)array_get_shape, NULL, "shape of the device array", NULL},
    {"device", (getter)array_get_device, NULL, "ordinal of the device holding the data", NULL},
    {"dtype", (getter)array_get_dtype, NULL, "numpy dtype of the device array", NULL},
    {NULL}};

static PyGetSetDef image_getset[] = {
    {"width", (getter)image_get_width, NULL, "image width in pixels", NULL},
    {"height", (getter)image_get_height, NULL, "image height in pixels", NULL},
    {NULL}};

static PyMethodDef array_methods[] = {
    {"__array__", (PyCFunction)array_array, METH_VARARGS | METH_KEYWORDS, "host copy as a numpy ndarray"},
    {"__array_ufunc__", (PyCFunction)array_ufunc, METH_VARARGS | METH_KEYWORDS, "numpy ufunc protocol (host)"},
    {"__array_function__", (PyCFunction)array_function, METH_VARARGS, "numpy function protocol (host)"},
    {"clone", (PyCFunction)array_clone, METH_NOARGS, "deep copy on the current target device"},
    {NULL}};

static PyMethodDef image_methods[] = {
    {"rgb2grey", (PyCFunction)image_grey, METH_NOARGS, "colour -> greyscale (luma 0.2125, 0.7154, 0.0721)"},
    {"rgb2gray", (PyCFunction)image_grey, METH_NOARGS, "alias of rgb2grey"},
    {"rgba2grey", (PyCFunction)image_grey, METH_NOARGS, "alias of rgb2grey"},
    {"rgba2gray", (PyCFunction)image_grey, METH_NOARGS, "alias of rgb2grey"},
    {"transpose", (PyCFunction)image_transpose, METH_NOARGS, "swap rows and columns (pixel is the unit)"},
    {"fliplr", (PyCFunction)image_fliplr, METH_NOARGS, "mirror every row"},
    {"rotate", (PyCFunction)image_rotate, METH_VARARGS, "rotate(angle_degrees) about the centre, same size"},
    {"random_rotate", (PyCFunction)image_rand_rotate, METH_VARARGS, "random_rotate(min, max)"},
    {"gaussian", (PyCFunction)image_gaussian, METH_VARARGS, "gaussian(sigma): separable blur, zero padding"},
    {"random_gaussian", (PyCFunction)image_rand_gaussian, METH_VARARGS, "random_gaussian(min, max)"},
    {"brightness", (PyCFunction)image_brightness, METH_VARARGS, "brightness(delta), |delta| < 1"},
    {"random_brightness", (PyCFunction)image_rand_brightness, METH_VARARGS, "random_brightness(min, max)"},
    {"adjust_gamma", (PyCFunction)image_gamma, METH_VARARGS, "adjust_gamma(gamma, gain)"},
    {"random_adjust_gamma", (PyCFunction)image_rand_gamma, METH_VARARGS,
     "random_adjust_gamma([gmin, gmax], [gainmin, gainmax])"},
    {"colorize", (PyCFunction)image_colorize, METH_VARARGS, "colorize(r, g, b) multipliers >= 0"},
    {"random_colorize", (PyCFunction)image_rand_colorize, METH_VARARGS,
     "random_colorize([r0, r1], [g0, g1], [b0, b1])"},
    {"clone", (PyCFunction)array_clone, METH_NOARGS, "deep copy on the current target device"},
    {NULL}};

PyTypeObject MPArray_Type = {
    PyVarObject_HEAD_INIT(NULL, 0).tp_name = "millipyde.gpuarray",
    .tp_basicsize = sizeof(MPArrayObject),
    .tp_dealloc = (destructor)array_dealloc,
    .tp_flags = Py_TPFLAGS_DEFAULT | Py_TPFLAGS_BASETYPE,
    .tp_doc = "An n-dimensional numeric array resident in GPU memory.",
    .tp_methods = array_methods,
    .tp_getset = array_getset,
    .tp_init = (initproc)array_init,
    .tp_new = array_new,
};

PyTypeObject MPImage_Type = {
    PyVarObject_HEAD_INIT(NULL, 0).tp_name = "millipyde.gpuimage",
    .tp_basicsize = sizeof(MPArrayObject),
    .tp_dealloc = (destructor)array_dealloc,
    .tp_flags = Py_TPFLAGS_DEFAULT | Py_TPFLAGS_BASETYPE,
    .tp_doc = "A 2-D (greyscale) or 3-D (H x W x channels) image resident in GPU memory.",
    .tp_methods = image_methods,
    .tp_getset = image_getset,
    .tp_init = (initproc)image_init,
    .tp_new = array_new,
};
