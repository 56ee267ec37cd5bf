/* quisk_b200/plugin/quisk_block_source.c -- boundary B4: a Quisk sample-source plugin.
 *
 * A Python extension module built the way the reference's own hardware plugins are (sdriqpkg/sdriq.c,
 * afedrinet/afedrinet_io.c, soapypkg/soapy.c): it imports the _quisk C API through the QUISK_C_API capsule
 * (import_quisk_api.c; quisk.h:441-466) and registers start / stop / read callbacks with
 * quisk_sample_source4 (sound.c:429-435).  quisk_read_sound (sound.c:938-939) then pulls its receive samples
 * from read_samples() below instead of a sound card or a UDP socket, and hands them to quisk_process_samples
 * (quisk.c:2289).
 *
 * The source is a block player: load(bytes, block) takes complex-double samples (CLIP32 scale, as every Quisk
 * source delivers them) and read hands out `block` of them per call until the buffer runs dry -- the synthetic
 * input of tests/test_quisk_swapin_gpu.py, and the shape of a source fed by a many-receiver front end.
 * It needs the reference's quisk.h to compile, so oracle/build_ref.sh builds it (from this file, against the
 * headers where they lie under /root/reference) into oracle/_ref/quisk_full/quisk_block_source.so.
 */
#include <Python.h>
#include <stdlib.h>
#include <string.h>
#include <complex.h>
#define IMPORT_QUISK_API
#include "quisk.h"

static complex double *samples_;
static long n_samples_, pos_;
static int block_ = 1024, started_, stopped_;

static void source_start(void) { started_++; }
static void source_stop(void) { stopped_++; }
static int source_read(complex double *cSamples)
{   /* ty_sample_read: fill cSamples (SAMP_BUFFER_SIZE entries available), return the count */
    long n = n_samples_ - pos_;
    if (n > block_) n = block_;
    if (n > SAMP_BUFFER_SIZE / 2) n = SAMP_BUFFER_SIZE / 2;
    if (n <= 0) return 0;
    memcpy(cSamples, samples_ + pos_, n * sizeof(complex double));
    pos_ += n;
    return (int)n;
}

static PyObject *load(PyObject *self, PyObject *args)
{
    Py_buffer view;
    int block;
    if (!PyArg_ParseTuple(args, "y*i", &view, &block)) return NULL;
    free(samples_);
    n_samples_ = view.len / (Py_ssize_t)sizeof(complex double);
    samples_ = (complex double *)malloc(view.len > 0 ? view.len : 1);
    memcpy(samples_, view.buf, view.len);
    PyBuffer_Release(&view);
    pos_ = 0; block_ = block > 0 ? block : 1024;
    return PyLong_FromLong(n_samples_);
}
static PyObject *open_samples(PyObject *self, PyObject *args)
{   /* what a hardware file's open() calls: register the callbacks with _quisk */
    if (!PyArg_ParseTuple(args, "")) return NULL;
    quisk_sample_source4(&source_start, &source_stop, &source_read, NULL);
    return PyUnicode_FromString("quisk_block_source: block player registered");
}
static PyObject *remaining(PyObject *self, PyObject *args) { return PyLong_FromLong(n_samples_ - pos_); }
static PyObject *status(PyObject *self, PyObject *args) { return Py_BuildValue("ii", started_, stopped_); }

static PyMethodDef methods[] = {
    {"load", load, METH_VARARGS, "load(bytes of complex128, block): samples to play, `block` per read."},
    {"open_samples", open_samples, METH_VARARGS, "Register start/stop/read with quisk_sample_source4."},
    {"remaining", remaining, METH_NOARGS, "Samples not yet handed out."},
    {"status", status, METH_NOARGS, "(start calls, stop calls)."},
    {NULL, NULL, 0, NULL}
};
static struct PyModuleDef moddef = { PyModuleDef_HEAD_INIT, "quisk_block_source", NULL, -1, methods };
PyMODINIT_FUNC PyInit_quisk_block_source(void)
{
    PyObject *m = PyModule_Create(&moddef);
    if (!m) return NULL;
    if (import_quisk_api()) { PyErr_SetString(PyExc_ImportError, "quisk_block_source: cannot import the _quisk C API"); return NULL; }
    return m;
}
