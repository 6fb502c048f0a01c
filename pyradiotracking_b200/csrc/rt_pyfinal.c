/*
 * rt_pyfinal.c -- CPython helper of the host finaliser: turns the column arrays of one engine call into `Signal` objects.
 *
 * Reference lines this replaces: the tail of extract_signals, radiotracking/analyze.py:434-450 -- one
 * `Signal(device, ts, frequency, duration, max, avg, std, noise, snr)` per detection, `ts = ts_start + timedelta(start_dt)` in UTC,
 * `duration = timedelta(seconds=duration_s)` -- which the Python loop of BatchAnalyzer.build_signals did at ~4.4 us per signal.  At
 * the rates a batched engine emits (configs[3]: ~3500 Signals per launch) that loop, not the GPU and not PCIe, bounded the
 * end-to-end throughput.  The arithmetic (microsecond rounding, float64 dB values) is done by the caller in numpy exactly as
 * before; this file only builds the objects: datetime + timedelta, timedelta, six floats, and either a call of the class or --
 * for the two known Signal classes, whose __init__ merely stores its arguments (radiotracking/__init__.py:136-170) -- a direct
 * fill of the instance dictionary with the same nine attributes.
 *
 * build_signals(cls, direct, devices, base_ts, unit, off_us, dur_us, freq, max, avg, std, noise, snr) -> list[list[Signal]]
 *   devices, base_ts : sequences with one entry per analyzer unit (device name; tz-aware start of the unit's block)
 *   unit, off_us, dur_us : int64 buffers; freq ... snr : float64 buffers; all of one length, rows in emission order
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <datetime.h>
#include <stdint.h>

static PyObject *k_device, *k_ts, *k_frequency, *k_duration, *k_max, *k_avg, *k_std, *k_noise, *k_snr;

static int get_buf(PyObject *o, Py_buffer *b, Py_ssize_t itemsize, Py_ssize_t n, const char *name) {
    if (PyObject_GetBuffer(o, b, PyBUF_CONTIG_RO) < 0) return -1;
    if (b->itemsize != itemsize || b->len != n * itemsize) {
        PyErr_Format(PyExc_ValueError, "%s: expected %zd contiguous items of %zd bytes", name, n, itemsize);
        PyBuffer_Release(b);
        return -1;
    }
    return 0;
}

static PyObject *delta_from_us(int64_t us) {
    int64_t days = us / 86400000000LL, rem = us % 86400000000LL;
    if (rem < 0) { rem += 86400000000LL; days -= 1; }
    return PyDelta_FromDSU((int)days, (int)(rem / 1000000), (int)(rem % 1000000));
}

static PyObject *build_signals(PyObject *self, PyObject *args) {
    PyObject *cls, *devices, *base_ts, *o[9];
    int direct;
    if (!PyArg_ParseTuple(args, "OpOOOOOOOOOOO", &cls, &direct, &devices, &base_ts, &o[0], &o[1], &o[2], &o[3], &o[4], &o[5], &o[6], &o[7], &o[8]))
        return NULL;
    if (!PyType_Check(cls)) { PyErr_SetString(PyExc_TypeError, "cls must be a class"); return NULL; }
    devices = PySequence_Fast(devices, "devices must be a sequence");
    if (!devices) return NULL;
    base_ts = PySequence_Fast(base_ts, "base_ts must be a sequence");
    if (!base_ts) { Py_DECREF(devices); return NULL; }
    const Py_ssize_t n_units = PySequence_Fast_GET_SIZE(devices);
    PyObject *out = NULL;
    Py_buffer b[9];
    int nb = 0;
    if (PySequence_Fast_GET_SIZE(base_ts) != n_units) { PyErr_SetString(PyExc_ValueError, "one base timestamp per unit"); goto done; }
    {
        Py_buffer first;
        if (PyObject_GetBuffer(o[0], &first, PyBUF_CONTIG_RO) < 0) goto done;
        const Py_ssize_t n = first.len / 8;
        PyBuffer_Release(&first);
        static const char *names[9] = {"unit", "off_us", "dur_us", "frequency", "max", "avg", "std", "noise", "snr"};
        for (; nb < 9; ++nb)
            if (get_buf(o[nb], &b[nb], 8, n, names[nb]) < 0) goto done;
        const int64_t *unit = (const int64_t *)b[0].buf, *off = (const int64_t *)b[1].buf, *dur = (const int64_t *)b[2].buf;
        const double *col[6];
        for (int c = 0; c < 6; ++c) col[c] = (const double *)b[3 + c].buf;
        PyObject *keys[6] = {k_frequency, k_max, k_avg, k_std, k_noise, k_snr};
        out = PyList_New(n_units);
        if (!out) goto done;
        for (Py_ssize_t u = 0; u < n_units; ++u) {
            PyObject *l = PyList_New(0);
            if (!l) { Py_CLEAR(out); goto done; }
            PyList_SET_ITEM(out, u, l);
        }
        PyObject *empty = PyTuple_New(0);
        if (!empty) { Py_CLEAR(out); goto done; }
        for (Py_ssize_t i = 0; i < n; ++i) {
            const int64_t u = unit[i];
            if (u < 0 || u >= n_units) { PyErr_SetString(PyExc_IndexError, "unit index out of range"); Py_CLEAR(out); break; }
            PyObject *d_off = delta_from_us(off[i]);
            PyObject *ts = d_off ? PyNumber_Add(PySequence_Fast_GET_ITEM(base_ts, u), d_off) : NULL;
            Py_XDECREF(d_off);
            PyObject *duration = delta_from_us(dur[i]);
            PyObject *f[6] = {NULL, NULL, NULL, NULL, NULL, NULL};
            int ok = ts && duration;
            for (int c = 0; c < 6 && ok; ++c) ok = (f[c] = PyFloat_FromDouble(col[c][i])) != NULL;
            PyObject *sig = NULL;
            if (ok && direct) {
                sig = ((PyTypeObject *)cls)->tp_new((PyTypeObject *)cls, empty, NULL);
                PyObject *d = sig ? PyObject_GenericGetDict(sig, NULL) : NULL;       /* the instance __dict__ (new reference) */
                if (d) {
                    ok = PyDict_SetItem(d, k_device, PySequence_Fast_GET_ITEM(devices, u)) == 0 && PyDict_SetItem(d, k_ts, ts) == 0;
                    ok = ok && PyDict_SetItem(d, keys[0], f[0]) == 0 && PyDict_SetItem(d, k_duration, duration) == 0;
                    for (int c = 1; c < 6 && ok; ++c) ok = PyDict_SetItem(d, keys[c], f[c]) == 0;
                    Py_DECREF(d);
                } else {
                    ok = 0;                                                             /* no instance dictionary: error is set */
                }
                if (!ok) Py_CLEAR(sig);
            } else if (ok) {
                PyObject *argv[9] = {PySequence_Fast_GET_ITEM(devices, u), ts, f[0], duration, f[1], f[2], f[3], f[4], f[5]};
                sig = PyObject_Vectorcall(cls, argv, 9, NULL);
            }
            Py_XDECREF(ts);
            Py_XDECREF(duration);
            for (int c = 0; c < 6; ++c) Py_XDECREF(f[c]);
            if (!sig || PyList_Append(PyList_GET_ITEM(out, u), sig) < 0) { Py_XDECREF(sig); Py_CLEAR(out); break; }
            Py_DECREF(sig);
        }
        Py_DECREF(empty);
    }
done:
    for (int i = 0; i < nb; ++i) PyBuffer_Release(&b[i]);
    Py_DECREF(devices);
    Py_DECREF(base_ts);
    return out;
}


/*
 * shadow_units(unit, off_us, dur_us, max, out) -> 0 | 1
 *   The shadow filter of SignalAnalyzer.filter_shadow_signals / is_shadow_of (radiotracking/analyze.py:282-328) for the signals of
 *   many analyzer units at once: rows sorted by unit; signal i is a shadow iff a signal j of the SAME unit overlaps it in time
 *   (closed intervals in integer microseconds: not ts_i > ts_j + dur_j, not ts_i + dur_i < ts_j) and is strictly louder.
 *   A unit holds the detections of one callback (tens of signals, ~800 on the loudest dense fixture), so every unit is an exact
 *   pairwise pass over its rows sorted by start time, cut off where the start times leave [ts_i - longest duration, ts_i + dur_i].
 *   Returns 1 without touching `out` if a unit holds more than 8192 signals (the caller then uses the O(S log S) numpy sweep).
 */
typedef struct { int64_t ts, te; double mx; Py_ssize_t row; } shadow_row;
static int shadow_cmp(const void *a, const void *b) {
    const shadow_row *x = (const shadow_row *)a, *y = (const shadow_row *)b;
    return x->ts < y->ts ? -1 : (x->ts > y->ts ? 1 : (x->row < y->row ? -1 : (x->row > y->row ? 1 : 0)));
}
static PyObject *shadow_units(PyObject *self, PyObject *args) {
    PyObject *o[5];
    if (!PyArg_ParseTuple(args, "OOOOO", &o[0], &o[1], &o[2], &o[3], &o[4])) return NULL;
    Py_buffer b[5];
    int nb = 0;
    PyObject *res = NULL;
    shadow_row *rows = NULL;
    Py_buffer first;
    if (PyObject_GetBuffer(o[0], &first, PyBUF_CONTIG_RO) < 0) return NULL;
    const Py_ssize_t n = first.len / 8;
    PyBuffer_Release(&first);
    static const char *names[4] = {"unit", "off_us", "dur_us", "max"};
    for (; nb < 4; ++nb)
        if (get_buf(o[nb], &b[nb], 8, n, names[nb]) < 0) goto done;
    if (PyObject_GetBuffer(o[4], &b[4], PyBUF_CONTIG) < 0) goto done;
    nb = 5;
    if (b[4].itemsize != 1 || b[4].len != n) { PyErr_SetString(PyExc_ValueError, "out: expected n bytes"); goto done; }
    {
        const int64_t *unit = (const int64_t *)b[0].buf, *off = (const int64_t *)b[1].buf, *dur = (const int64_t *)b[2].buf;
        const double *mx = (const double *)b[3].buf;
        unsigned char *out = (unsigned char *)b[4].buf;
        Py_ssize_t longest = 0;
        for (Py_ssize_t i = 0, j; i < n; i = j) {
            for (j = i + 1; j < n && unit[j] == unit[i]; ++j) {}
            if (j - i > longest) longest = j - i;
        }
        if (longest > 8192) { res = PyLong_FromLong(1); goto done; }
        rows = (shadow_row *)PyMem_Malloc((size_t)(longest > 0 ? longest : 1) * sizeof(shadow_row));
        if (!rows) { PyErr_NoMemory(); goto done; }
        for (Py_ssize_t i = 0, j; i < n; i = j) {
            for (j = i + 1; j < n && unit[j] == unit[i]; ++j) {}
            const Py_ssize_t k = j - i;
            int64_t dmax = 0;
            for (Py_ssize_t r = 0; r < k; ++r) {
                rows[r].ts = off[i + r]; rows[r].te = off[i + r] + dur[i + r]; rows[r].mx = mx[i + r]; rows[r].row = i + r;
                if (dur[i + r] > dmax) dmax = dur[i + r];
            }
            qsort(rows, (size_t)k, sizeof(shadow_row), shadow_cmp);
            for (Py_ssize_t r = 0; r < k; ++r) {
                unsigned char sh = 0;
                /* later starts: while ts_q <= te_r (they overlap iff te_q >= ts_r, which holds since ts_q >= ts_r) */
                for (Py_ssize_t q = r + 1; q < k && rows[q].ts <= rows[r].te; ++q)
                    if (rows[q].mx > rows[r].mx) { sh = 1; break; }
                /* earlier or equal starts: overlap iff te_q >= ts_r; none can reach ts_r once ts_q < ts_r - dmax */
                for (Py_ssize_t q = r - 1; !sh && q >= 0 && rows[q].ts >= rows[r].ts - dmax; --q)
                    if (rows[q].te >= rows[r].ts && rows[q].mx > rows[r].mx) sh = 1;
                out[rows[r].row] = sh;
            }
        }
        res = PyLong_FromLong(0);
    }
done:
    PyMem_Free(rows);
    for (int i = 0; i < nb; ++i) PyBuffer_Release(&b[i]);
    return res;
}

static PyMethodDef methods[] = {
    {"shadow_units", shadow_units, METH_VARARGS, "shadow filter (analyze.py:282-328) per analyzer unit on column arrays"},
    {"build_signals", build_signals, METH_VARARGS, "column arrays of one engine call -> per-unit lists of Signal objects"},
    {NULL, NULL, 0, NULL},
};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_rtfinal", "Signal object builder of the host finaliser", -1, methods};

PyMODINIT_FUNC PyInit__rtfinal(void) {
    PyDateTime_IMPORT;
    if (!PyDateTimeAPI) return NULL;
    k_device = PyUnicode_InternFromString("device");
    k_ts = PyUnicode_InternFromString("ts");
    k_frequency = PyUnicode_InternFromString("frequency");
    k_duration = PyUnicode_InternFromString("duration");
    k_max = PyUnicode_InternFromString("max");
    k_avg = PyUnicode_InternFromString("avg");
    k_std = PyUnicode_InternFromString("std");
    k_noise = PyUnicode_InternFromString("noise");
    k_snr = PyUnicode_InternFromString("snr");
    return PyModule_Create(&moddef);
}
