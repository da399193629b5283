/*
 * Operation, Pipeline, Generator, Device.
 *
 * Reference: src/gpuoperation.c (ctor :45-122, run/run_on :126-185, coin :187-210,
 * name tables :214-291), src/gpupipeline.c (ctor :57-170, connect_to :186-218,
 * run :234-312), src/gpugenerator.c (ctor :45-170, iterator :173-281),
 * src/device.c (:28-70).  Validation rules and error strings are the
 * reference's.  Pipeline.run() and the Generator hand the whole chain to the C
 * executor (include/mp_pipeline.h) with the GIL released.
 */
#include "ext_common.h"

/* ==================================================================== Operation */
static void operation_dealloc(MPOperationObject *self)
{
    Py_XDECREF(self->callable);
    Py_XDECREF(self->arg_tuple);
    Py_TYPE(self)->tp_free((PyObject *)self);
}

static PyObject *operation_new(PyTypeObject *type, PyObject *args, PyObject *kwds)
{
    MPOperationObject *self = (MPOperationObject *)type->tp_alloc(type, 0);
    if (self) {
        self->callable = NULL;
        self->arg_tuple = NULL;
        self->requires_instance = 0;
        self->probability = -1.0;
    }
    return (PyObject *)self;
}

static int operation_init(MPOperationObject *self, PyObject *args, PyObject *kwds)
{
    const Py_ssize_t nargs = PyTuple_Size(args);
    if (nargs < 1) {
        PyErr_SetString(PyExc_ValueError, mperr_str(GPUOPERATION_ERROR_CONSTRUCTION_NO_ARGS));
        return -1;
    }
    self->probability = -1.0;
    if (kwds && PyDict_Size(kwds) > 0) {
        PyObject *prob = PyDict_GetItemString(kwds, "probability");
        if (!prob || PyDict_Size(kwds) != 1) {
            PyErr_SetString(PyExc_ValueError, mperr_str(GPUOPERATION_ERROR_CONSTRUCTION_NAMED_ARGS));
            return -1;
        }
        /* a Python float strictly between 0 and 1 (tests/millipyde_tests.py:140-174) */
        if (!PyFloat_Check(prob)) {
            PyErr_SetString(PyExc_ValueError, mperr_str(GPUOPERATION_ERROR_INVALID_PROBABILITY));
            return -1;
        }
        const double p = PyFloat_AsDouble(prob);
        if (p <= 0.0 || p >= 1.0) {
            PyErr_SetString(PyExc_ValueError, mperr_str(GPUOPERATION_ERROR_INVALID_PROBABILITY));
            return -1;
        }
        self->probability = p;
    }
    PyObject *callable = PyTuple_GetItem(args, 0);
    self->requires_instance = PyUnicode_Check(callable) ? 1 : 0;
    PyObject *call_args = PyTuple_GetSlice(args, 1, nargs);
    if (!call_args) return -1;
    Py_INCREF(callable);
    Py_XSETREF(self->callable, callable);
    Py_XSETREF(self->arg_tuple, call_args);
    return 0;
}

static int coin_says_run(double probability, int *run)
{
    *run = 1;
    if (probability < 0) return 0;
    double u;
    MPStatus st = random_double_in_range(0.0, 1.0, &u);
    if (st != MILLIPYDE_SUCCESS) {
        mpext_raise_status(st, "Operation");
        return -1;
    }
    if (u > probability) *run = 0;
    return 0;
}

PyObject *mpext_operation_run(MPOperationObject *op)
{
    int run;
    if (coin_says_run(op->probability, &run) < 0) return NULL;
    if (!run) Py_RETURN_NONE;
    return PyObject_Call(op->callable, op->arg_tuple, NULL);
}

PyObject *mpext_operation_run_on(MPOperationObject *op, PyObject *instance)
{
    if (!PyUnicode_Check(op->callable)) {
        PyErr_SetString(PyExc_ValueError, mperr_str(GPUOPERATION_ERROR_RUN_WITHOUT_STRING_METHOD));
        return NULL;
    }
    PyObject *method = PyObject_GetAttr(instance, op->callable);
    if (!method) {
        PyErr_Clear();
        PyErr_SetString(PyExc_ValueError, mperr_str(GPUOPERATION_ERROR_RUN_UNKNOWN_STRING_METHOD));
        return NULL;
    }
    int run;
    PyObject *res = NULL;
    if (coin_says_run(op->probability, &run) == 0) {
        if (run) {
            res = PyObject_Call(method, op->arg_tuple, NULL);
        } else {
            res = Py_None;
            Py_INCREF(res);
        }
    }
    Py_DECREF(method);
    return res;
}

static PyObject *operation_run(MPOperationObject *self, PyObject *Py_UNUSED(ignored))
{
    return mpext_operation_run(self);
}
static PyObject *operation_run_on(MPOperationObject *self, PyObject *instance)
{
    return mpext_operation_run_on(self, instance);
}

static int list_pair(PyObject *o, double *lo, double *hi)
{
    if (!PyList_Check(o) || PyList_Size(o) != 2) {
        PyErr_SetString(PyExc_ValueError, "expected a [min, max] list");
        return -1;
    }
    *lo = PyFloat_AsDouble(PyList_GetItem(o, 0));
    *hi = PyFloat_AsDouble(PyList_GetItem(o, 1));
    return PyErr_Occurred() ? -1 : 0;
}

/* gpuoperation_func_from_name + gpuoperation_args_from_name (src/gpuoperation.c:214-291),
 * over the library's name table (which also knows the grey spellings and random_*). */
int mpext_resolve_operation(MPOperationObject *op, MPRunnable *out)
{
    memset(out, 0, sizeof *out);
    out->probability = op->probability;
    if (!op->requires_instance) return 0;
    const char *name = PyUnicode_AsUTF8(op->callable);
    if (!name) return -1;
    size_t bytes = 0;
    MPFunc fn = mpimg_func_from_name(name, &bytes);
    if (!fn) return 0;
    out->func = fn;
    if (!bytes) {
        if (PyTuple_Size(op->arg_tuple) != 0) {
            PyErr_Format(PyExc_TypeError, "%s() takes no arguments", name);
            return -1;
        }
        return 1;
    }
    double *a = (double *)calloc(1, bytes);
    int ok = 0;
    PyObject *t = op->arg_tuple;
    if (fn == mpimg_rotate || fn == mpimg_gaussian) {
        ok = PyArg_ParseTuple(t, "d", &a[0]);
    } else if (fn == mpimg_brightness) {
        ok = PyArg_ParseTuple(t, "d", &a[0]);
        if (ok && (a[0] <= -1 || a[0] >= 1)) {
            PyErr_SetString(PyExc_ValueError, "brightness delta must lie strictly between -1 and 1");
            ok = 0;
        }
    } else if (fn == mpimg_adjust_gamma) {
        ok = PyArg_ParseTuple(t, "dd", &a[0], &a[1]);
    } else if (fn == mpimg_colorize) {
        ok = PyArg_ParseTuple(t, "ddd", &a[0], &a[1], &a[2]);
        if (ok && (a[0] < 0 || a[1] < 0 || a[2] < 0)) {
            PyErr_SetString(PyExc_ValueError, "colorize multipliers must be >= 0");
            ok = 0;
        }
    } else if (fn == mpimg_random_rotate || fn == mpimg_random_gaussian || fn == mpimg_random_brightness) {
        ok = PyArg_ParseTuple(t, "dd", &a[0], &a[1]);
    } else if (fn == mpimg_random_adjust_gamma) {
        PyObject *g, *k;
        ok = PyArg_ParseTuple(t, "OO", &g, &k) && list_pair(g, &a[0], &a[1]) == 0 && list_pair(k, &a[2], &a[3]) == 0;
    } else if (fn == mpimg_random_colorize) {
        PyObject *r, *g, *b;
        ok = PyArg_ParseTuple(t, "OOO", &r, &g, &b) && list_pair(r, &a[0], &a[1]) == 0 &&
             list_pair(g, &a[2], &a[3]) == 0 && list_pair(b, &a[4], &a[5]) == 0;
    }
    if (!ok) {
        free(a);
        return -1;
    }
    out->args = a;
    return 1;
}

static PyMethodDef operation_methods[] = {
    {"run", (PyCFunction)operation_run, METH_NOARGS, "call the wrapped callable with the stored arguments"},
    {"run_on", (PyCFunction)operation_run_on, METH_O, "call the named method on the given object"},
    {NULL}};

static PyMemberDef operation_members[] = {
    {"probability", T_DOUBLE, offsetof(MPOperationObject, probability), READONLY, "-1 when unconditional"},
    {NULL}};

PyTypeObject MPOperation_Type = {
    PyVarObject_HEAD_INIT(NULL, 0).tp_name = "millipyde.Operation",
    .tp_basicsize = sizeof(MPOperationObject),
    .tp_dealloc = (destructor)operation_dealloc,
    .tp_flags = Py_TPFLAGS_DEFAULT,
    .tp_doc = "Operation(callable_or_method_name, *args, probability=p)",
    .tp_methods = operation_methods,
    .tp_members = operation_members,
    .tp_init = (initproc)operation_init,
    .tp_new = operation_new,
};

/* ===================================================================== Pipeline */
/* Resolve a list of Operations into an MPPipeline.  Operations that are not
 * string-named C operators become NULL stages, which the executor skips exactly
 * as the reference does (src/gpupipeline.c:393-396).  *all_resolved reports
 * whether every one resolved. */
static MPPipeline *build_pipe(PyObject *operations, int device_id, int *all_resolved)
{
    const Py_ssize_t n = PyList_Size(operations);
    MPRunnable *stages = (MPRunnable *)calloc(n > 0 ? (size_t)n : 1, sizeof(MPRunnable));
    int failed = 0;
    *all_resolved = 1;
    for (Py_ssize_t i = 0; i < n && !failed; ++i) {
        PyObject *op = PyList_GetItem(operations, i);
        if (!PyObject_TypeCheck(op, &MPOperation_Type)) {
            PyErr_SetString(PyExc_ValueError, mperr_str(GPUPIPELINE_ERROR_NONLIST_OPERATIONS));
            failed = 1;
            break;
        }
        int r = mpext_resolve_operation((MPOperationObject *)op, &stages[i]);
        if (r < 0) failed = 1;
        if (r == 0) *all_resolved = 0;
    }
    MPPipeline *pipe = failed ? NULL : mppipe_create(stages, (int)n, device_id);
    for (Py_ssize_t i = 0; i < n; ++i) free(stages[i].args); /* mppipe_create copied them */
    free(stages);
    if (!pipe && !failed) PyErr_NoMemory();
    return pipe;
}

static void pipeline_dealloc(MPPipelineObject *self)
{
    Py_XDECREF(self->inputs);
    Py_XDECREF(self->operations);
    Py_XDECREF((PyObject *)self->receiver);
    if (self->pipe) mppipe_destroy(self->pipe);
    Py_TYPE(self)->tp_free((PyObject *)self);
}

static PyObject *pipeline_new(PyTypeObject *type, PyObject *args, PyObject *kwds)
{
    MPPipelineObject *self = (MPPipelineObject *)type->tp_alloc(type, 0);
    if (self) {
        self->inputs = self->operations = NULL;
        self->pipe = NULL;
        self->receiver = NULL;
    }
    return (PyObject *)self;
}

static int pipeline_init(MPPipelineObject *self, PyObject *args, PyObject *kwds)
{
    const Py_ssize_t nargs = PyTuple_Size(args);
    if (nargs < 2 || nargs > 3) {
        PyErr_SetString(PyExc_ValueError, mperr_str(GPUPIPELINE_ERROR_CONSTRUCTION_INVALID_ARGS));
        return -1;
    }
    PyObject *inputs = PyTuple_GetItem(args, 0), *operations = PyTuple_GetItem(args, 1);
    if (!PyList_CheckExact(inputs)) {
        PyErr_SetString(PyExc_ValueError, mperr_str(GPUPIPELINE_ERROR_NONLIST_INPUTS));
        return -1;
    }
    if (!PyList_CheckExact(operations)) {
        PyErr_SetString(PyExc_ValueError, mperr_str(GPUPIPELINE_ERROR_NONLIST_OPERATIONS));
        return -1;
    }
    int device_id;
    if (kwds && PyDict_Size(kwds) > 0) {
        PyObject *dev = PyDict_GetItemString(kwds, "device");
        if (!dev || PyDict_Size(kwds) != 1) {
            PyErr_SetString(PyExc_ValueError, mperr_str(GPUPIPELINE_ERROR_CONSTRUCTION_NAMED_ARGS));
            return -1;
        }
        if (!PyLong_Check(dev)) {
            PyErr_SetString(PyExc_ValueError, mperr_str(GPUPIPELINE_ERROR_INVALID_DEVICE));
            return -1;
        }
        device_id = (int)PyLong_AsLong(dev);
        if (!mpdev_is_valid_device(device_id)) {
            PyErr_SetString(PyExc_ValueError, mperr_str(GPUPIPELINE_ERROR_UNUSABLE_DEVICE));
            return -1;
        }
    } else {
        device_id = mpdev_get_target_device(); /* possibly "no affinity": spread over all devices at run() */
    }
    for (Py_ssize_t i = 0; i < PyList_Size(inputs); ++i) {
        if (!MP_IS_GPU_OBJECT(PyList_GetItem(inputs, i))) {
            PyErr_SetString(PyExc_ValueError, mperr_str(GPUPIPELINE_ERROR_NONGPU_INPUT));
            return -1;
        }
    }
    if (self->pipe) {
        /* a second __init__: another Pipeline may hold this one's executor as its receiver
         * (connect_to), so it cannot be replaced underneath it */
        PyErr_SetString(PyExc_RuntimeError, "Pipeline is already initialised; construct a new one");
        return -1;
    }
    int all;
    MPPipeline *pipe = build_pipe(operations, device_id, &all);
    if (!pipe) return -1;
    self->pipe = pipe;
    Py_INCREF(inputs);
    Py_INCREF(operations);
    Py_XSETREF(self->inputs, inputs);
    Py_XSETREF(self->operations, operations);
    return 0;
}

static PyObject *pipeline_connect_to(MPPipelineObject *self, PyObject *other)
{
    if (!PyObject_TypeCheck(other, &MPPipeline_Type)) {
        PyErr_SetString(PyExc_TypeError, "connect_to() expects a Pipeline");
        return NULL;
    }
    MPPipelineObject *recv = (MPPipelineObject *)other;
    mppipe_connect(self->pipe, recv->pipe); /* device auto-assignment rules of the reference */
    Py_INCREF(other);
    Py_XSETREF(self->receiver, recv);
    Py_RETURN_NONE;
}

static PyObject *pipeline_run(MPPipelineObject *self, PyObject *Py_UNUSED(ignored))
{
    if (mpext_require_devices() < 0) return NULL;
    const Py_ssize_t n = PyList_Size(self->inputs);
    MPObjData **objs = (MPObjData **)calloc(n > 0 ? (size_t)n : 1, sizeof(MPObjData *));
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject *item = PyList_GetItem(self->inputs, i);
        if (!MP_IS_GPU_OBJECT(item) || !((MPArrayObject *)item)->obj) {
            free(objs);
            PyErr_SetString(PyExc_ValueError, mperr_str(GPUPIPELINE_ERROR_NONGPU_INPUT));
            return NULL;
        }
        objs[i] = ((MPArrayObject *)item)->obj;
    }
    MPStatus st;
    Py_BEGIN_ALLOW_THREADS
    st = mppipe_run(self->pipe, objs, (int)n);
    Py_END_ALLOW_THREADS
    free(objs);
    if (st != MILLIPYDE_SUCCESS) return mpext_raise_status(st, "Pipeline.run");
    Py_RETURN_NONE;
}

static PyObject *pipeline_get_device(MPPipelineObject *self, void *closure)
{
    return PyLong_FromLong(self->pipe ? mppipe_get_device(self->pipe) : DEVICE_LOC_NO_AFFINITY);
}
static PyObject *pipeline_get_launches(MPPipelineObject *self, void *closure)
{
    return PyLong_FromUnsignedLongLong(self->pipe ? mppipe_last_launches(self->pipe) : 0);
}
static PyObject *pipeline_get_segments(MPPipelineObject *self, void *closure)
{
    return PyLong_FromLong(self->pipe ? mppipe_last_segments(self->pipe) : 0);
}

static PyMethodDef pipeline_methods[] = {
    {"run", (PyCFunction)pipeline_run, METH_NOARGS, "run every operation on every input; returns when all devices are idle"},
    {"connect_to", (PyCFunction)pipeline_connect_to, METH_O, "feed this pipeline's results into another one"},
    {NULL}};

static PyGetSetDef pipeline_getset[] = {
    {"device", (getter)pipeline_get_device, NULL, "device the pipeline is bound to (-2: none)", NULL},
    {"last_launches", (getter)pipeline_get_launches, NULL, "kernel launches of the most recent run()", NULL},
    {"last_segments", (getter)pipeline_get_segments, NULL, "fused segments of the most recent run()", NULL},
    {NULL}};

PyTypeObject MPPipeline_Type = {
    PyVarObject_HEAD_INIT(NULL, 0).tp_name = "millipyde.Pipeline",
    .tp_basicsize = sizeof(MPPipelineObject),
    .tp_dealloc = (destructor)pipeline_dealloc,
    .tp_flags = Py_TPFLAGS_DEFAULT,
    .tp_doc = "Pipeline(inputs: list, operations: list, device=int)",
    .tp_methods = pipeline_methods,
    .tp_getset = pipeline_getset,
    .tp_init = (initproc)pipeline_init,
    .tp_new = pipeline_new,
};

/* ==================================================================== Generator */
#define NO_OUTPUT_MAX (-1L)

static void generator_dealloc(MPGeneratorObject *self)
{
    if (self->inflight && self->pipe) { /* the devices may still be reading what the batch borrows */
        Py_BEGIN_ALLOW_THREADS
        mppipe_wait(self->pipe);
        Py_END_ALLOW_THREADS
    }
    Py_XDECREF(self->inflight);
    Py_XDECREF(self->inflight_keep);
    Py_XDECREF(self->inputs);
    Py_XDECREF(self->operations);
    Py_XDECREF(self->ready);
    if (self->pipe) mppipe_destroy(self->pipe);
    Py_TYPE(self)->tp_free((PyObject *)self);
}

static PyObject *generator_new(PyTypeObject *type, PyObject *args, PyObject *kwds)
{
    MPGeneratorObject *self = (MPGeneratorObject *)type->tp_alloc(type, 0);
    if (self) {
        self->inputs = self->operations = self->ready = NULL;
        self->pipe = NULL;
        self->device_id = DEVICE_LOC_NO_AFFINITY;
        self->max = NO_OUTPUT_MAX;
        self->produced = self->i = 0;
        self->return_to_host = 0;
        self->prefetch = 0;
        self->inflight = self->inflight_keep = NULL;
        self->ready_pos = 0;
    }
    return (PyObject *)self;
}

static int generator_init(MPGeneratorObject *self, PyObject *args, PyObject *kwds)
{
    const Py_ssize_t nargs = PyTuple_Size(args);
    if (nargs < 2 || nargs > 5) {
        PyErr_SetString(PyExc_ValueError, mperr_str(GPUGENERATOR_ERROR_INVALID_INPUT));
        return -1;
    }
    PyObject *inputs = PyTuple_GetItem(args, 0), *operations = PyTuple_GetItem(args, 1);
    if (kwds) {
        PyObject *key, *value;
        Py_ssize_t pos = 0;
        while (PyDict_Next(kwds, &pos, &key, &value)) {
            const char *k = PyUnicode_AsUTF8(key);
            if (!k || (strcmp(k, "device") && strcmp(k, "outputs") && strcmp(k, "return_to_host") &&
                       strcmp(k, "prefetch"))) {
                PyErr_SetString(PyExc_ValueError, mperr_str(GPUGENERATOR_ERROR_CONSTRUCTION_NAMED_ARGS));
                return -1;
            }
        }
        PyObject *dev = PyDict_GetItemString(kwds, "device");
        PyObject *outs = PyDict_GetItemString(kwds, "outputs");
        PyObject *ret = PyDict_GetItemString(kwds, "return_to_host");
        PyObject *pre = PyDict_GetItemString(kwds, "prefetch");
        if (dev) {
            if (!PyLong_Check(dev)) {
                PyErr_SetString(PyExc_ValueError, mperr_str(GPUGENERATOR_ERROR_INVALID_DEVICE));
                return -1;
            }
            self->device_id = (int)PyLong_AsLong(dev);
            if (!mpdev_is_valid_device(self->device_id)) {
                PyErr_SetString(PyExc_ValueError, mperr_str(GPUGENERATOR_ERROR_UNUSABLE_DEVICE));
                return -1;
            }
        }
        if (outs) {
            if (!PyLong_Check(outs) || PyLong_AsLong(outs) < 0) {
                PyErr_SetString(PyExc_ValueError, mperr_str(GPUGENERATOR_ERROR_INVALID_MAX));
                return -1;
            }
            self->max = PyLong_AsLong(outs);
        }
        if (ret) {
            if (!PyBool_Check(ret)) {
                PyErr_SetString(PyExc_ValueError, mperr_str(GPUGENERATOR_ERROR_INVALID_RETURN_TO));
                return -1;
            }
            self->return_to_host = PyObject_IsTrue(ret);
        }
        if (pre) {
            if (!PyLong_Check(pre) || PyLong_AsLong(pre) < 0) {
                PyErr_SetString(PyExc_ValueError, "prefetch must be a non-negative integer");
                return -1;
            }
            self->prefetch = (int)PyLong_AsLong(pre);
        }
    }
    if (!PyList_CheckExact(inputs) && !PyUnicode_Check(inputs)) {
        PyErr_SetString(PyExc_ValueError, mperr_str(GPUGENERATOR_ERROR_INVALID_INPUT));
        return -1;
    }
    if (!PyList_CheckExact(operations)) {
        PyErr_SetString(PyExc_ValueError, mperr_str(GPUGENERATOR_ERROR_NONLIST_OPERATIONS));
        return -1;
    }
    PyObject *in_list;
    if (PyUnicode_Check(inputs)) {
        in_list = mpext_images_from_path(inputs);
        if (!in_list) return -1;
    } else {
        in_list = inputs;
        Py_INCREF(in_list);
    }
    for (Py_ssize_t i = 0; i < PyList_Size(in_list); ++i) {
        if (!MP_IS_GPU_OBJECT(PyList_GetItem(in_list, i))) {
            Py_DECREF(in_list);
            PyErr_SetString(PyExc_ValueError, mperr_str(TYPE_ERROR_NON_GPUOBJ));
            return -1;
        }
    }
    /* If every op is a string-named C operator the whole chain runs in the C executor (fused,
     * batched, multi-device); otherwise items are produced one by one through Python calls like
     * the reference does (src/gpugenerator.c:246-270). */
    int all = 1;
    MPPipeline *pipe = NULL;
    for (Py_ssize_t i = 0; i < PyList_Size(operations); ++i)
        if (!PyObject_TypeCheck(PyList_GetItem(operations, i), &MPOperation_Type)) {
            Py_DECREF(in_list);
            PyErr_SetString(PyExc_ValueError, mperr_str(GPUGENERATOR_ERROR_NONLIST_OPERATIONS));
            return -1;
        }
    pipe = build_pipe(operations, self->device_id, &all);
    if (!pipe) {
        Py_DECREF(in_list);
        return -1;
    }
    if (!all) {
        mppipe_destroy(pipe);
        pipe = NULL;
    }
    if (self->pipe) mppipe_destroy(self->pipe);
    self->pipe = pipe;
    Py_XSETREF(self->inputs, in_list);
    Py_INCREF(operations);
    Py_XSETREF(self->operations, operations);
    Py_XSETREF(self->ready, PyList_New(0));
    self->produced = self->i = 0;
    return 0;
}

static int generator_device(MPGeneratorObject *g)
{
    int dev = g->device_id;
    if (dev == DEVICE_LOC_NO_AFFINITY) dev = mpdev_get_target_device();
    return dev; /* may still be "no affinity": the batch path then spreads over every device */
}

/* Reference path: clone -> ops through Python attribute lookup -> optional D2H. */
static PyObject *produce_one_python(MPGeneratorObject *g, long index)
{
    const Py_ssize_t n_in = PyList_Size(g->inputs);
    PyObject *input = PyList_GetItem(g->inputs, index % n_in);
    int dev = generator_device(g);
    if (dev == DEVICE_LOC_NO_AFFINITY) dev = mpdev_get_recommended_device();
    PyObject *result = mpext_clone((MPArrayObject *)input, dev, 0);
    if (!result) return NULL;
    for (Py_ssize_t k = 0; k < PyList_Size(g->operations); ++k) {
        MPOperationObject *op = (MPOperationObject *)PyList_GetItem(g->operations, k);
        PyObject *r = op->requires_instance ? mpext_operation_run_on(op, result) : mpext_operation_run(op);
        if (!r) {
            Py_DECREF(result);
            return NULL;
        }
        Py_DECREF(r);
    }
    if (g->return_to_host) {
        PyObject *host = mpext_to_ndarray((MPArrayObject *)result);
        Py_DECREF(result);
        return host;
    }
    return result;
}

/* Executor path: clone the next `count` inputs (spread over the devices when no device is bound),
 * run the chain on all of them in one mppipe_run (GIL released), append to the ready list. */
/* ---- page-locked result arrays of Generator(return_to_host=True) --------------------------
 * The reference downloads every output with a blocking pageable copy (src/gpugenerator.c:273-279).
 * Here the outputs of a batch are downloaded together, asynchronously, into page-locked ndarrays
 * (full PCIe rate, one wait per batch).  The page-locked blocks are recycled: the ndarray's base
 * object is a capsule whose destructor hands the block back to a small cache, so a steady stream
 * of same-sized outputs allocates nothing. */
#define MP_PIN_CACHE_SLOTS 1024
#define MP_PIN_CACHE_BYTES ((size_t)8 << 30) /* a look-ahead of 256 outputs of 12.6 MB is 3.2 GB */
typedef struct {
    void *p;
    size_t bytes;
} PinBlock;
static PinBlock g_pin_cache[MP_PIN_CACHE_SLOTS];
static size_t g_pin_cached = 0;

static void *pin_take(size_t bytes)
{
    for (int i = 0; i < MP_PIN_CACHE_SLOTS; ++i)
        if (g_pin_cache[i].p && g_pin_cache[i].bytes == bytes) {
            void *p = g_pin_cache[i].p;
            g_pin_cache[i].p = NULL;
            g_pin_cached -= bytes;
            return p;
        }
    return mphost_alloc_pinned(bytes);
}

static void pin_capsule_free(PyObject *capsule)
{
    PinBlock *b = (PinBlock *)PyCapsule_GetPointer(capsule, "mp_pinned_result");
    if (!b) return;
    if (g_pin_cached + b->bytes <= MP_PIN_CACHE_BYTES)
        for (int i = 0; i < MP_PIN_CACHE_SLOTS; ++i)
            if (!g_pin_cache[i].p) {
                g_pin_cache[i] = *b;
                g_pin_cached += b->bytes;
                free(b);
                return;
            }
    mphost_free_pinned(b->p);
    free(b);
}

/* ndarray shaped like `o` over a page-locked block; NULL (no exception) if none can be had. */
static PyObject *pinned_result_array(const MPObjData *o)
{
    if (!o->nbytes) return NULL;
    PinBlock *b = (PinBlock *)malloc(sizeof(PinBlock));
    if (!b) return NULL;
    b->bytes = o->nbytes;
    b->p = pin_take(o->nbytes);
    if (!b->p) {
        free(b);
        return NULL;
    }
    npy_intp dims[NPY_MAXDIMS];
    for (int i = 0; i < o->ndims; ++i) dims[i] = o->dims[i];
    PyObject *arr = PyArray_SimpleNewFromData(o->ndims, dims, o->type, b->p);
    PyObject *cap = arr ? PyCapsule_New(b, "mp_pinned_result", pin_capsule_free) : NULL;
    if (!cap || PyArray_SetBaseObject((PyArrayObject *)arr, cap) < 0) { /* SetBaseObject steals cap */
        PyErr_Clear();
        if (!cap) {
            mphost_free_pinned(b->p);
            free(b);
        }
        Py_XDECREF(arr);
        return NULL;
    }
    return arr;
}

/* A look-ahead this long or longer (explicit prefetch=) runs asynchronously: batch k + 1 is on the
 * devices while the consumer drains batch k.  Its views then borrow from per-batch REPLICAS of the
 * inputs (one device-side copy of each input per device per batch: 6 / prefetch of the stream's
 * traffic), so nothing the caller does to an input while iterating can reach a batch in flight. */
#define MP_ASYNC_PREFETCH 32

/* Build the next `count` outputs as views / clones and hand them to the executor.  sync != 0: wait and
 * leave the finished batch in g->inflight; else return with the batch running. */
static int submit_batch(MPGeneratorObject *g, long count, int async_mode)
{
    const Py_ssize_t n_in = PyList_Size(g->inputs);
    const int bound = generator_device(g);
    const int ndev = mpdev_get_device_count();
    PyObject *batch = PyList_New(0);
    PyObject *keep = PyList_New(0);
    MPObjData **objs = (MPObjData **)calloc((size_t)count, sizeof(MPObjData *));
    int dev = bound != DEVICE_LOC_NO_AFFINITY ? bound : mpdev_get_recommended_device();
    /* One device and every input already there: hand the executor VIEWS of the inputs instead of
     * clones.  The chain's first launch then reads the input itself and writes the output's own
     * buffer, which saves the clone's pass over every image (mppipe_run_views). */
    int use_views = !async_mode && (bound != DEVICE_LOC_NO_AFFINITY || ndev == 1);
    for (Py_ssize_t k = 0; use_views && k < n_in; ++k) {
        MPObjData *o = ((MPArrayObject *)PyList_GetItem(g->inputs, k))->obj;
        if (!o || !o->device_data || o->mem_loc != dev) use_views = 0;
    }
    /* Spreading over several devices, or running ahead of the consumer: every device gets ONE replica of
     * each input it needs for this batch (a device-side / peer copy made now, so a caller who mutated an
     * input since the last batch is honoured and one who mutates it during this one cannot hurt) and its
     * outputs are views of that replica -- not one clone per output. */
    const int spread = bound == DEVICE_LOC_NO_AFFINITY && ndev > 1;
    const int replicate = !use_views && (spread || async_mode);
    PyObject **replica = replicate ? (PyObject **)calloc((size_t)ndev * (size_t)n_in, sizeof(PyObject *)) : NULL;
    int rep_ok = replicate && replica != NULL && batch && keep && objs;
    for (Py_ssize_t k = 0; rep_ok && k < n_in; ++k) {
        MPObjData *o = ((MPArrayObject *)PyList_GetItem(g->inputs, k))->obj;
        if (!o || !o->device_data) rep_ok = 0;
    }
    int failed = !batch || !keep || !objs;
    long made = 0;
    for (long k = 0; !failed && k < count; ++k) {
        const Py_ssize_t which = (g->produced + k) % n_in;
        PyObject *input = PyList_GetItem(g->inputs, which);
        PyObject *c;
        if (use_views) {
            c = mpext_wrap_obj(Py_TYPE(input), mpobj_view_data(((MPArrayObject *)input)->obj));
        } else if (rep_ok) {
            PyObject **slot = &replica[(size_t)dev * (size_t)n_in + (size_t)which];
            if (!*slot) {
                if (!async_mode && ((MPArrayObject *)input)->obj->mem_loc == dev) {
                    Py_INCREF(input);
                    *slot = input;
                } else {
                    *slot = mpext_clone((MPArrayObject *)input, dev, 0);
                }
                if (*slot) PyList_Append(keep, *slot);
            }
            c = *slot ? mpext_wrap_obj(Py_TYPE(input), mpobj_view_data(((MPArrayObject *)*slot)->obj)) : NULL;
        } else {
            /* the executor assigns image k to device block k / THREADS_PER_DEVICE: clone it there */
            c = mpext_clone((MPArrayObject *)input, dev, 0);
        }
        if (!c) {
            failed = 1;
            break;
        }
        objs[k] = ((MPArrayObject *)c)->obj;
        PyList_Append(batch, c);
        Py_DECREF(c);
        ++made;
        if (spread && (k + 1) % THREADS_PER_DEVICE == 0) dev = mpdev_get_next_device(dev);
    }
    if (replica) {
        for (size_t r = 0; r < (size_t)ndev * (size_t)n_in; ++r) Py_XDECREF(replica[r]); /* `keep` holds them */
        free(replica);
    }
    const int borrowing = use_views || rep_ok;
    MPStatus st = MILLIPYDE_SUCCESS;
    if (!failed) {
        /* one random-source key for the Generator's whole stream, images numbered by output index: output k
         * is the same draw whatever the look-ahead and however the batch is spread over the devices */
        mppipe_hold_run_key(g->pipe);
        mppipe_set_index_base(g->pipe, (unsigned long long)g->produced);
        Py_BEGIN_ALLOW_THREADS
        if (borrowing) st = async_mode ? mppipe_submit_views(g->pipe, objs, (int)count) : mppipe_run_views(g->pipe, objs, (int)count);
        else st = async_mode ? mppipe_submit(g->pipe, objs, (int)count) : mppipe_run(g->pipe, objs, (int)count);
        Py_END_ALLOW_THREADS
    } else if (borrowing) { /* the views made so far still borrow: they must not free what they borrow */
        for (long j = 0; j < made; ++j) objs[j]->device_data = NULL;
    }
    free(objs);
    if (failed || st != MILLIPYDE_SUCCESS) {
        Py_XDECREF(batch);
        Py_XDECREF(keep);
        if (!failed) mpext_raise_status(st, "Generator");
        else if (!PyErr_Occurred()) PyErr_NoMemory();
        return -1;
    }
    Py_XSETREF(g->inflight, batch);
    Py_XSETREF(g->inflight_keep, keep);
    g->produced += count;
    return 0;
}

/* Wait for the batch in flight (if it was submitted asynchronously), download it if the Generator
 * returns host arrays, and append it to the ready list. */
static int collect_batch(MPGeneratorObject *g, int was_async)
{
    PyObject *batch = g->inflight;
    if (!batch) return 0;
    g->inflight = NULL;
    const long count = (long)PyList_Size(batch);
    MPStatus st = MILLIPYDE_SUCCESS;
    if (was_async) {
        Py_BEGIN_ALLOW_THREADS
        st = mppipe_wait(g->pipe);
        Py_END_ALLOW_THREADS
    }
    Py_CLEAR(g->inflight_keep); /* every view owns its result now; the replicas retire in stream order */
    if (st != MILLIPYDE_SUCCESS) {
        Py_DECREF(batch);
        mpext_raise_status(st, "Generator");
        return -1;
    }
    if (g->ready_pos > 0) { /* drop what has been handed out */
        PyList_SetSlice(g->ready, 0, g->ready_pos, NULL);
        g->ready_pos = 0;
    }
    if (g->return_to_host) {
        /* all downloads of the batch in flight at once, into page-locked arrays, one wait each */
        PyObject **hosts = (PyObject **)calloc((size_t)count, sizeof(PyObject *));
        MPObjData **res = (MPObjData **)calloc((size_t)count, sizeof(MPObjData *));
        if (!hosts || !res) {
            free(hosts);
            free(res);
            Py_DECREF(batch);
            PyErr_NoMemory();
            return -1;
        }
        for (long k = 0; k < count; ++k) {
            res[k] = ((MPArrayObject *)PyList_GetItem(batch, k))->obj;
            hosts[k] = pinned_result_array(res[k]);
        }
        Py_BEGIN_ALLOW_THREADS
        for (long k = 0; k < count && st == MILLIPYDE_SUCCESS; ++k)
            if (hosts[k]) st = mpobj_download_async(res[k], PyArray_DATA((PyArrayObject *)hosts[k]), res[k]->nbytes);
        for (long k = 0; k < count; ++k)
            if (hosts[k]) {
                MPStatus s2 = mpobj_synchronize(res[k]);
                if (st == MILLIPYDE_SUCCESS) st = s2;
            }
        Py_END_ALLOW_THREADS
        int failed = st != MILLIPYDE_SUCCESS;
        if (failed) mpext_raise_status(st, "Generator");
        for (long k = 0; k < count; ++k) {
            if (!failed && !hosts[k]) { /* no page-locked block to be had: the plain blocking copy */
                hosts[k] = mpext_to_ndarray((MPArrayObject *)PyList_GetItem(batch, k));
                if (!hosts[k]) failed = 1;
            }
            if (!failed) PyList_Append(g->ready, hosts[k]);
            Py_XDECREF(hosts[k]);
        }
        free(hosts);
        free(res);
        if (failed) {
            Py_DECREF(batch);
            return -1;
        }
    } else {
        for (long k = 0; k < count; ++k) PyList_Append(g->ready, PyList_GetItem(batch, k));
    }
    Py_DECREF(batch);
    return 0;
}

static PyObject *generator_iter(PyObject *self)
{
    Py_INCREF(self);
    return self;
}

static PyObject *generator_next(MPGeneratorObject *g)
{
    if (!g->inputs || !g->operations) return NULL; /* exhausted earlier */
    if (g->max != NO_OUTPUT_MAX && g->i >= g->max) {
        Py_CLEAR(g->inputs);
        Py_CLEAR(g->operations);
        Py_CLEAR(g->ready);
        return NULL; /* StopIteration */
    }
    if (PyList_Size(g->inputs) == 0) {
        PyErr_SetString(PyExc_ValueError, mperr_str(GPUGENERATOR_ERROR_INVALID_INPUT));
        return NULL;
    }
    if (mpext_require_devices() < 0) return NULL;
    if (!g->pipe) {
        PyObject *r = produce_one_python(g, g->i);
        if (r) g->i++;
        return r;
    }
    if (g->ready_pos >= PyList_Size(g->ready)) {
        /* default look-ahead: one block of THREADS_PER_DEVICE items per device when spreading, two
         * blocks on a single device -- enough for the executor to batch and overlap, small enough
         * that a consumer that stops early has paid for at most a few items it never sees.  An explicit
         * look-ahead of MP_ASYNC_PREFETCH or more runs one batch ahead of the consumer. */
        long want = g->prefetch;
        if (want <= 0)
            want = (generator_device(g) == DEVICE_LOC_NO_AFFINITY && mpdev_get_device_count() > 1)
                       ? (long)THREADS_PER_DEVICE * mpdev_get_device_count()
                       : 2L * THREADS_PER_DEVICE;
        const int async_mode = g->prefetch >= MP_ASYNC_PREFETCH;
        long next = want;
        if (g->max != NO_OUTPUT_MAX && g->produced + next > g->max) next = g->max - g->produced;
        if (!g->inflight) { /* nothing running yet (first call, or the synchronous mode) */
            if (next < 1) next = 1;
            if (submit_batch(g, next, async_mode) < 0) return NULL;
        }
        if (collect_batch(g, async_mode) < 0) return NULL;
        if (async_mode) { /* start the batch after this one: it runs while the consumer iterates */
            next = want;
            if (g->max != NO_OUTPUT_MAX && g->produced + next > g->max) next = g->max - g->produced;
            if (next >= 1 && submit_batch(g, next, 1) < 0) return NULL;
        }
    }
    PyObject *item = PyList_GetItem(g->ready, g->ready_pos);
    if (!item) return NULL;
    Py_INCREF(item);
    Py_INCREF(Py_None);
    PyList_SetItem(g->ready, g->ready_pos, Py_None); /* hand the reference over: the list keeps no output alive */
    g->ready_pos++;
    g->i++;
    return item;
}

PyTypeObject MPGenerator_Type = {
    PyVarObject_HEAD_INIT(NULL, 0).tp_name = "millipyde.Generator",
    .tp_basicsize = sizeof(MPGeneratorObject),
    .tp_dealloc = (destructor)generator_dealloc,
    .tp_flags = Py_TPFLAGS_DEFAULT,
    .tp_doc = "Generator(inputs: list | dir_path, operations: list, device=int, outputs=int, "
              "return_to_host=bool, prefetch=int)",
    .tp_iter = generator_iter,
    .tp_iternext = (iternextfunc)generator_next,
    .tp_init = (initproc)generator_init,
    .tp_new = generator_new,
};

/* ======================================================================= Device */
static PyObject *device_new(PyTypeObject *type, PyObject *args, PyObject *kwds)
{
    MPDeviceObject *self = (MPDeviceObject *)type->tp_alloc(type, 0);
    if (self) self->device_id = self->prev_device_id = DEVICE_LOC_NO_AFFINITY;
    return (PyObject *)self;
}

static int device_init(MPDeviceObject *self, PyObject *args, PyObject *kwds)
{
    int id;
    if (!PyArg_ParseTuple(args, "i", &id)) return -1;
    self->device_id = id;
    return 0;
}

static PyObject *device_enter(MPDeviceObject *self, PyObject *Py_UNUSED(ignored))
{
    if (!mpdev_is_valid_device(self->device_id)) {
        PyErr_Format(PyExc_ValueError, "device %d is not usable", self->device_id);
        return NULL;
    }
    self->prev_device_id = mpdev_get_target_device();
    mpdev_set_target_device(self->device_id);
    return PyLong_FromLong(self->device_id);
}

/* On an exception: drain the device, clear its error state, restore the target and let the
 * exception propagate (src/device.c:60-65 resets the whole device, which would free every live
 * gpuimage; see mpdev_reset).  Otherwise synchronise and restore. */
static PyObject *device_exit(MPDeviceObject *self, PyObject *args)
{
    PyObject *et, *ev, *tb;
    if (!PyArg_ParseTuple(args, "OOO", &et, &ev, &tb)) return NULL;
    if (et != Py_None) {
        mpdev_reset(self->device_id);
        mpdev_set_target_device(self->prev_device_id);
        Py_RETURN_FALSE;
    }
    Py_BEGIN_ALLOW_THREADS
    mpdev_set_device(self->device_id);
    mpdev_synchronize();
    Py_END_ALLOW_THREADS
    mpdev_set_target_device(self->prev_device_id);
    Py_RETURN_FALSE;
}

static PyMethodDef device_methods[] = {
    {"__enter__", (PyCFunction)device_enter, METH_NOARGS, "make this the target device"},
    {"__exit__", (PyCFunction)device_exit, METH_VARARGS, "synchronise and restore the previous target"},
    {NULL}};

static PyMemberDef device_members[] = {
    {"id", T_INT, offsetof(MPDeviceObject, device_id), READONLY, "device ordinal"},
    {NULL}};

PyTypeObject MPDevice_Type = {
    PyVarObject_HEAD_INIT(NULL, 0).tp_name = "millipyde.Device",
    .tp_basicsize = sizeof(MPDeviceObject),
    .tp_flags = Py_TPFLAGS_DEFAULT,
    .tp_doc = "Device(id): context manager that sets the target device",
    .tp_methods = device_methods,
    .tp_members = device_members,
    .tp_init = (initproc)device_init,
    .tp_new = device_new,
};
