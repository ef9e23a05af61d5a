"""CPU tests of the Python mirror's host logic (no GPU call): how list-valued occupation arguments are
turned into (input, output) job pairs -- the reference's list branch,
ChinHuhPermanentCalculator_Wrapper.hpp:233-326 -- and which `lib` ids are refused."""
import numpy as np
import pytest

from piquassoboost_b200.Boson_Sampling_Utilities import (ChinHuhPermanentCalculator_wrapper as W,
                                                         GlynnPermanentCalculator_wrapper as G)

a, b, c = np.array([1, 0, 1]), np.array([0, 2, 0]), np.array([1, 1, 0])


def test_reference_form_input_list_with_per_input_output_lists():
    w = W(7, np.eye(3), [a, b], [[a, c], [b]])
    pairs, shape = w._pairs()
    assert shape == [2, 1]
    assert [(p[0].tolist(), p[1].tolist()) for p in pairs] == [
        (a.tolist(), a.tolist()), (a.tolist(), c.tolist()), (b.tolist(), b.tolist())]


def test_reference_form_multi_input():
    """input_state[i] is the list of inputs that go with output_state[i] (multiInput branch)."""
    w = W(7, np.eye(3), [[a, c], [b]], [a, b])
    pairs, shape = w._pairs()
    assert shape == [2, 1]
    assert [(p[0].tolist(), p[1].tolist()) for p in pairs] == [
        (a.tolist(), a.tolist()), (c.tolist(), a.tolist()), (b.tolist(), b.tolist())]


def test_flat_lists_are_the_cross_product_extension():
    w = W(7, np.eye(3), [a, b], [a, b, c])
    pairs, (ins, outs) = w._pairs()
    assert pairs is None and len(ins) == 2 and len(outs) == 3


def test_ragged_lengths_must_match():
    with pytest.raises(Exception):
        W(7, np.eye(3), [a, b], [[a]])._pairs()


def test_single_vectors_stay_vectors():
    w = W(7, np.eye(3), [1, 0, 1], (0, 2, 0))
    assert w.input_state.tolist() == [1, 0, 1] and w.output_state.tolist() == [0, 2, 0]


def test_unsupported_lib_ids_raise_like_a_reference_build_without_them():
    with pytest.raises(Exception, match="DFE"):
        W(2, np.eye(3), a, a).calculate()
    with pytest.raises(Exception, match="Wrong value"):
        W(77, np.eye(3), a, a).calculate()
    with pytest.raises(Exception, match="DFE"):
        G(np.eye(3), 3).calculate()
    with pytest.raises(Exception, match="Wrong value"):
        G(np.eye(3), 9).calculate()
