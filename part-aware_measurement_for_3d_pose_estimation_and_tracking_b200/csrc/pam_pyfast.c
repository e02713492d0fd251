/* pam_pyfast.c -- CPython helper of the per-frame drop-in path (dropin/IterativeTracker.tracking).
 *
 * The reference's caller hands tracking() a python list with one (m, J, 3) float64 array per camera
 * (src/ivclabpose.py:216-257).  Packing that list into the mapped slot of the resident kernel with numpy costs
 * 10-15 us of interpreter time per frame -- as much as the frame itself takes on the device.  This module does the
 * packing (float64 -> float32 with a representability check), the submit and the wait in ONE call through the buffer
 * protocol.  It holds no algorithmic code: the frame is processed by libpam.so's resident kernel; the functions it
 * calls (pam_stream_submit / pam_stream_wait, include/pam.h) are passed in as addresses by the loader.
 *
 * Build: gcc -O2 -shared -fPIC -I<python include> (see __graft_entry__.build()). */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>

typedef int (*submit_fn)(void*, int32_t);
typedef int (*wait_fn)(void*);
static submit_fn g_submit = NULL;
static wait_fn g_wait = NULL;

static PyObject* bind(PyObject* self, PyObject* args) {
    unsigned long long a, b;
    if (!PyArg_ParseTuple(args, "KK", &a, &b)) return NULL;
    g_submit = (submit_fn)(uintptr_t)a;
    g_wait = (wait_fn)(uintptr_t)b;
    Py_RETURN_NONE;
}

/* track_frame(handle, frame_id, detections_list, packed_addr, counts_addr, V, D, row_floats, strict)
 *   -> (rc, flags)   rc: libpam status of submit/wait;  flags: bit 0 = a camera had more than D detections
 *                    (the first D were used), bit 1 = a value was not float32-representable (strict: the frame was
 *                    NOT submitted and rc = 0)
 * Raises TypeError when an element of the list is not a C-contiguous float64 / float32 buffer of shape (m, J, 3)
 * (the caller then falls back to its numpy path). */
static PyObject* track_frame(PyObject* self, PyObject* args) {
    unsigned long long handle, packed_addr, counts_addr;
    int frame_id, V, D, row, strict;
    PyObject* lst;
    if (!PyArg_ParseTuple(args, "KiOKKiiii", &handle, &frame_id, &lst, &packed_addr, &counts_addr, &V, &D, &row, &strict))
        return NULL;
    if (!g_submit || !g_wait) { PyErr_SetString(PyExc_RuntimeError, "pam_pyfast.bind() has not been called"); return NULL; }
    PyObject* seq = PySequence_Fast(lst, "detections_list must be a sequence");
    if (!seq) return NULL;
    if (PySequence_Fast_GET_SIZE(seq) != V) {
        Py_DECREF(seq);
        PyErr_SetString(PyExc_TypeError, "one entry per camera expected");
        return NULL;
    }
    float* packed = (float*)(uintptr_t)packed_addr;
    int32_t* counts = (int32_t*)(uintptr_t)counts_addr;
    int flags = 0;
    Py_ssize_t rows = 0;
    for (int c = 0; c < V; ++c) {
        PyObject* it = PySequence_Fast_GET_ITEM(seq, c);
        Py_buffer vw;
        if (PyObject_GetBuffer(it, &vw, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) != 0) {
            PyErr_Clear();
            Py_DECREF(seq);
            PyErr_SetString(PyExc_TypeError, "not a contiguous buffer");
            return NULL;
        }
        Py_ssize_t m = 0;
        int ok = 1, is_f64 = 0;
        if (vw.len == 0) m = 0;                                   /* np.array([]) of a camera without detections */
        else if (vw.ndim == 3 && vw.shape[1] * vw.shape[2] == row && vw.format &&
                 ((vw.format[0] == 'd' && vw.itemsize == 8) || (vw.format[0] == 'f' && vw.itemsize == 4))) {
            m = vw.shape[0];
            is_f64 = vw.format[0] == 'd';
        } else ok = 0;
        if (!ok) {
            PyBuffer_Release(&vw);
            Py_DECREF(seq);
            PyErr_SetString(PyExc_TypeError, "expected a (m, J, 3) float64 / float32 array per camera");
            return NULL;
        }
        if (m > D) { m = D; flags |= 1; }
        float* dst = packed + rows * row;
        const Py_ssize_t n = m * row;
        if (is_f64) {
            const double* src = (const double*)vw.buf;
            int bad = 0;
            for (Py_ssize_t k = 0; k < n; ++k) {
                const float f = (float)src[k];
                dst[k] = f;
                bad |= ((double)f != src[k]) & (src[k] == src[k]);   /* NaN stays NaN: not a rounding */
            }
            if (bad) flags |= 2;
        } else if (n) {
            memcpy(dst, vw.buf, (size_t)n * 4);
        }
        counts[c] = (int32_t)m;
        rows += m;
        PyBuffer_Release(&vw);
    }
    Py_DECREF(seq);
    int rc = 0;
    if (!((flags & 2) && strict)) {
        rc = g_submit((void*)(uintptr_t)handle, (int32_t)frame_id);
        if (rc == 0) {
            Py_BEGIN_ALLOW_THREADS
            rc = g_wait((void*)(uintptr_t)handle);
            Py_END_ALLOW_THREADS
        }
    }
    return Py_BuildValue("ii", rc, flags);
}

static PyMethodDef methods[] = {
    {"bind", bind, METH_VARARGS, "bind(submit_addr, wait_addr): addresses of pam_stream_submit / pam_stream_wait"},
    {"track_frame", track_frame, METH_VARARGS, "pack one frame into the mapped slot, submit it, wait for the results"},
    {NULL, NULL, 0, NULL}};
static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "pam_pyfast", "per-frame drop-in helper of libpam.so", -1, methods};
PyMODINIT_FUNC PyInit_pam_pyfast(void) { return PyModule_Create(&moddef); }
