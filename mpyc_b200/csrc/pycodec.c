/* _pycodec: Python ints <-> fixed-width little-endian limb buffers at C speed.
 *
 * MPyC hands field arrays over as NumPy dtype=object arrays of Python ints (mpyc/finfields.py:703-725);
 * the kernels work on 8*L-byte little-endian elements.  Converting through int.to_bytes / int.from_bytes
 * in Python costs 0.2-0.4 us per element and dominated the drop-in path; these two functions do it in
 * ~30 ns per element with the CPython long API.
 *
 *   pack(seq, nbytes, modulus) -> bytearray of len(seq)*nbytes     (values reduced mod modulus if needed)
 *   unpack(buffer, nbytes)     -> list of ints
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <string.h>

static int as_bytes(PyObject* v, unsigned char* dst, size_t n) {
#if PY_VERSION_HEX >= 0x030D0000
    return _PyLong_AsByteArray((PyLongObject*)v, dst, n, 1, 0, 1);
#else
    return _PyLong_AsByteArray((PyLongObject*)v, dst, n, 1, 0);
#endif
}

static PyObject* pack(PyObject* self, PyObject* args) {
    PyObject *seq, *modulus;
    Py_ssize_t nbytes;
    if (!PyArg_ParseTuple(args, "OnO", &seq, &nbytes, &modulus)) return NULL;
    if (!PyLong_Check(modulus) || nbytes < 1) {
        PyErr_SetString(PyExc_TypeError, "pack(seq, nbytes, modulus:int)");
        return NULL;
    }
    PyObject* fast = PySequence_Fast(seq, "pack: expected a sequence of ints");
    if (!fast) return NULL;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
    PyObject* out = PyByteArray_FromStringAndSize(NULL, n * nbytes);
    if (!out) {
        Py_DECREF(fast);
        return NULL;
    }
    unsigned char* dst = (unsigned char*)PyByteArray_AS_STRING(out);
    const size_t kbits = _PyLong_NumBits(modulus);
    PyObject** items = PySequence_Fast_ITEMS(fast);
    for (Py_ssize_t i = 0; i < n; i++) {
        PyObject* v = items[i];
        PyObject* owned = NULL;
        if (!PyLong_Check(v)) {            /* field elements / numpy ints / polynomials: go through int() */
            owned = PyNumber_Long(v);
            if (!owned) goto fail;
            v = owned;
        }
        int reduce = 0;
        if (_PyLong_Sign(v) < 0) reduce = 1;
        else {
            size_t bits = _PyLong_NumBits(v);
            if (bits > kbits) reduce = 1;
            else if (bits == kbits) {
                int lt = PyObject_RichCompareBool(v, modulus, Py_LT);
                if (lt < 0) { Py_XDECREF(owned); goto fail; }
                reduce = !lt;
            }
        }
        if (reduce) {
            PyObject* r = PyNumber_Remainder(v, modulus);
            Py_XDECREF(owned);
            if (!r) goto fail;
            owned = r;
            v = r;
        }
        int rc = as_bytes(v, dst + i * nbytes, (size_t)nbytes);
        Py_XDECREF(owned);
        if (rc < 0) goto fail;
    }
    Py_DECREF(fast);
    return out;
fail:
    Py_DECREF(fast);
    Py_DECREF(out);
    return NULL;
}

static PyObject* unpack(PyObject* self, PyObject* args) {
    Py_buffer buf;
    Py_ssize_t nbytes;
    if (!PyArg_ParseTuple(args, "y*n", &buf, &nbytes)) return NULL;
    if (nbytes < 1 || buf.len % nbytes) {
        PyBuffer_Release(&buf);
        PyErr_SetString(PyExc_ValueError, "unpack: buffer length is not a multiple of nbytes");
        return NULL;
    }
    const Py_ssize_t n = buf.len / nbytes;
    PyObject* out = PyList_New(n);
    if (!out) {
        PyBuffer_Release(&buf);
        return NULL;
    }
    const unsigned char* src = (const unsigned char*)buf.buf;
    for (Py_ssize_t i = 0; i < n; i++) {
        PyObject* v = _PyLong_FromByteArray(src + i * nbytes, (size_t)nbytes, 1, 0);
        if (!v) {
            Py_DECREF(out);
            PyBuffer_Release(&buf);
            return NULL;
        }
        PyList_SET_ITEM(out, i, v);
    }
    PyBuffer_Release(&buf);
    return out;
}

static PyMethodDef methods[] = {
    {"pack", pack, METH_VARARGS, "pack(seq, nbytes, modulus) -> bytearray (little-endian, reduced mod modulus)"},
    {"unpack", unpack, METH_VARARGS, "unpack(buffer, nbytes) -> list of ints"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "_pycodec", "int <-> limb buffer codec", -1, methods};

PyMODINIT_FUNC PyInit__pycodec(void) { return PyModule_Create(&moduledef); }
