// oracle/dropin/Boson_Sampling_Utilities_wrapper.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Module initialisation for the reference's CPython permanent wrappers compiled UNMODIFIED from
// /root/reference/piquassoboost/sampling/{ChinHuh,Glynn}PermanentCalculator_Wrapper.hpp against
// include/dropin/ (calculator classes backed by libpqb200.so).  The reference's own module file
// (sampling/Boson_Sampling_Utilities_wrapper.cpp:29-127) also registers four hafnian wrapper types, which
// are outside the hot path of this repository (out of scope): they are registered
// here as placeholders that raise, so that the reference's Boson_Sampling_Utilities.py imports unchanged.
#define PY_SSIZE_T_CLEAN

#include <Python.h>
#include <numpy/arrayobject.h>
#include "ChinHuhPermanentCalculator_Wrapper.hpp"
#include "GlynnPermanentCalculator_Wrapper.hpp"

extern "C" {

static int placeholder_init(PyObject*, PyObject*, PyObject*) {
    PyErr_SetString(PyExc_NotImplementedError, "hafnian calculators are outside the scope of the B200 permanent engine");
    return -1;
}

static PyTypeObject make_placeholder(const char* name) {
    PyTypeObject t = {PyVarObject_HEAD_INIT(NULL, 0)};
    t.tp_name = name;
    t.tp_basicsize = sizeof(PyObject);
    t.tp_flags = Py_TPFLAGS_DEFAULT | Py_TPFLAGS_BASETYPE;
    t.tp_new = PyType_GenericNew;
    t.tp_init = placeholder_init;
    return t;
}

static PyTypeObject g_placeholders[4];
static const char* g_placeholder_names[4] = {"PowerTraceHafnian_wrapper", "PowerTraceHafnianRecursive_wrapper",
                                             "PowerTraceLoopHafnian_wrapper", "PowerTraceLoopHafnianRecursive_wrapper"};
static const char* g_placeholder_qualnames[4] = {
    "Boson_Sampling_Utilities_wrapper.PowerTraceHafnian_wrapper",
    "Boson_Sampling_Utilities_wrapper.PowerTraceHafnianRecursive_wrapper",
    "Boson_Sampling_Utilities_wrapper.PowerTraceLoopHafnian_wrapper",
    "Boson_Sampling_Utilities_wrapper.PowerTraceLoopHafnianRecursive_wrapper"};

static PyModuleDef Boson_Sampling_Utilities_wrapper_Module = {
    PyModuleDef_HEAD_INIT, "Boson_Sampling_Utilities_wrapper",
    "the reference's permanent wrapper types over the B200 engine (test build)", -1,
};

PyMODINIT_FUNC PyInit_Boson_Sampling_Utilities_wrapper(void) {
    import_array();
    if (PyType_Ready(&ChinHuhPermanentCalculator_wrapper_Type) < 0) return NULL;
    if (PyType_Ready(&GlynnPermanentCalculator_wrapper_Type) < 0) return NULL;
    PyObject* m = PyModule_Create(&Boson_Sampling_Utilities_wrapper_Module);
    if (m == NULL) return NULL;
    Py_INCREF(&ChinHuhPermanentCalculator_wrapper_Type);
    if (PyModule_AddObject(m, "ChinHuhPermanentCalculator_wrapper", (PyObject*)&ChinHuhPermanentCalculator_wrapper_Type) < 0) return NULL;
    Py_INCREF(&GlynnPermanentCalculator_wrapper_Type);
    if (PyModule_AddObject(m, "GlynnPermanentCalculator_wrapper", (PyObject*)&GlynnPermanentCalculator_wrapper_Type) < 0) return NULL;
    for (int k = 0; k < 4; k++) {
        g_placeholders[k] = make_placeholder(g_placeholder_qualnames[k]);
        if (PyType_Ready(&g_placeholders[k]) < 0) return NULL;
        Py_INCREF(&g_placeholders[k]);
        if (PyModule_AddObject(m, g_placeholder_names[k], (PyObject*)&g_placeholders[k]) < 0) return NULL;
    }
    return m;
}

}  // extern "C"
