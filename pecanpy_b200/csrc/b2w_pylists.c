/* b2w_pylists.c -- the hand-off of a walk matrix to Python: uint32[rows, L+2] -> List[List[str]].
 *
 * Replaces the per-walk Python comprehension Base._map_walk (reference pecanpy.py:103-114, called for every
 * row at pecanpy.py:160): `[self.nodes[i] for i in walk_idx_ary[:end_idx]]` with end_idx = last column.
 * After the GPU kernel this mapping is what dominates simulate_walks, and it cannot
 * leave the CPU -- the result is made of Python objects -- so it is one tight C loop over the matrix:
 * one PyList per row, filled by pointer copies out of the id list (no per-element Python bytecode, no
 * intermediate object array), with the id objects prefetched a few entries ahead (the id table of a
 * 10^6-node graph does not fit the CPU caches and the indices of a walk are random).
 *
 * A CPython extension (not ctypes) because it creates Python objects; built by pecanpy_b200/build.py with
 * gcc next to libb2w.so.  Holds the GIL throughout (reference counts).
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

/* rows_to_lists(matrix, ids, walk_length, row_begin, row_end) -> list of lists
 *   matrix: C-contiguous-row buffer of 4-byte unsigned integers, shape (rows, >= walk_length + 2)
 *   ids:    list (or tuple) of node id objects; row entry v maps to ids[v]
 * Row r yields ids[matrix[r, k]] for k < matrix[r, walk_length + 1] (the effective length). */
static PyObject* rows_to_lists(PyObject* self, PyObject* args) {
  PyObject *mat_obj, *ids_obj;
  Py_ssize_t L, r0, r1;
  (void)self;
  if (!PyArg_ParseTuple(args, "OOnnn", &mat_obj, &ids_obj, &L, &r0, &r1)) return NULL;
  PyObject* ids_fast = PySequence_Fast(ids_obj, "ids must be a sequence");
  if (!ids_fast) return NULL;
  Py_buffer view;
  if (PyObject_GetBuffer(mat_obj, &view, PyBUF_STRIDES | PyBUF_FORMAT) != 0) { Py_DECREF(ids_fast); return NULL; }
  PyObject* out = NULL;
  if (view.ndim != 2 || view.itemsize != 4 || view.strides[1] != 4 || L < 1 || view.shape[1] < L + 2) {
    PyErr_SetString(PyExc_ValueError, "matrix must be a 2-D array of 4-byte integers with unit column stride and "
                                      "at least walk_length + 2 columns");
    goto done;
  }
  if (r0 < 0 || r1 > view.shape[0] || r0 > r1) { PyErr_SetString(PyExc_IndexError, "row range out of bounds"); goto done; }
  {
    const Py_ssize_t n_ids = PySequence_Fast_GET_SIZE(ids_fast);
    PyObject** items = PySequence_Fast_ITEMS(ids_fast);
    const char* base = (const char*)view.buf;
    const Py_ssize_t rs = view.strides[0];
    out = PyList_New(r1 - r0);
    if (!out) goto done;
    for (Py_ssize_t r = r0; r < r1; ++r) {
      const uint32_t* row = (const uint32_t*)(base + r * rs);
      const Py_ssize_t len = (Py_ssize_t)row[L + 1];
      if (len > L + 1) {
        PyErr_Format(PyExc_ValueError, "row %zd: effective length %zd exceeds walk_length + 1", r, len);
        Py_CLEAR(out);
        goto done;
      }
      PyObject* lst = PyList_New(len);
      if (!lst) { Py_CLEAR(out); goto done; }
      PyList_SET_ITEM(out, r - r0, lst);
      for (Py_ssize_t k = 0; k < len; ++k) {
        /* two-level prefetch: the slot of the id table 16 entries ahead, the id object itself 8 ahead */
        if (k + 16 < len && (Py_ssize_t)row[k + 16] < n_ids) __builtin_prefetch(&items[row[k + 16]], 0, 1);
        if (k + 8 < len && (Py_ssize_t)row[k + 8] < n_ids) __builtin_prefetch(items[row[k + 8]], 1, 1);
        const Py_ssize_t v = (Py_ssize_t)row[k];
        if (v >= n_ids) {
          PyErr_Format(PyExc_IndexError, "row %zd: node index %zd out of range for %zd ids", r, v, n_ids);
          /* the partially filled list holds NULL slots; fill them so that deallocation is safe */
          for (Py_ssize_t t = k; t < len; ++t) { Py_INCREF(Py_None); PyList_SET_ITEM(lst, t, Py_None); }
          Py_CLEAR(out);
          goto done;
        }
        PyObject* o = items[v];
        Py_INCREF(o);
        PyList_SET_ITEM(lst, k, o);
      }
    }
  }
done:
  PyBuffer_Release(&view);
  Py_DECREF(ids_fast);
  return out;
}

static PyMethodDef methods[] = {
    {"rows_to_lists", rows_to_lists, METH_VARARGS,
     "rows_to_lists(matrix, ids, walk_length, row_begin, row_end) -> List[List[id]] (pecanpy.py:103-114 for a block of rows)"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "_b2w_pylists",
                                       "walk matrix -> Python lists (hand-off after the B200 walk kernel)", -1, methods,
                                       NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__b2w_pylists(void) { return PyModule_Create(&moduledef); }
