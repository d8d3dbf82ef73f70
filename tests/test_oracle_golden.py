"""Pins the oracle against every golden vector the reference's tests hold for this path:
material matrices only.  Element/global quantities have no golden values in the reference; they are pinned
by the reference's property tests restated in test_oracle_properties.py."""
import numpy as np

import oracle_lib as O


def _D(E, nu):
    D = np.zeros(36)
    O.lib().oracle_stiffness_3d(E, nu, D.ctypes.data)
    return D.reshape(6, 6)


def test_stiffness_3d_golden(built):
    # /root/reference/tests/material/mast_isotropic_material.cpp:684-694 and
    # tests/property/mast_3d_isotropic_element_property_card.cpp:422-433  (E=72e9, nu=0.33)
    D = _D(72e9, 0.33)
    ref = np.zeros((6, 6))
    ref[:3, :3] = 0.525431225121628e11
    for i in range(3):
        ref[i, i] = 1.066784608580274e11
        ref[i + 3, i + 3] = 0.541353383458647e11 / 2.0
    assert np.allclose(D, ref, rtol=1e-14, atol=0)


def test_thermal_expansion_A_golden(built):
    # tests/property/mast_3d_isotropic_element_property_card.cpp:207-210: (D*alpha)_i = 1.149882352941177e7
    # for E=72e9, nu=0.33, alpha=5.43e-5
    v = np.zeros(6)
    O.lib().oracle_thermal_expansion_A_3d(72e9, 0.33, 5.43e-5, v.ctypes.data)
    assert np.allclose(v[:3], 1.149882352941177e7, rtol=1e-13)
    assert np.all(v[3:] == 0.0)


def test_stiffness_derivative_matches_fd(built):
    # analogue of tests/material/mast_isotropic_material_sensitivity.cpp: dD/dE and dD/dnu
    E, nu = 72e9, 0.33
    d = np.zeros(36)
    O.lib().oracle_stiffness_3d_derivative(E, nu, 1.0, 0.0, d.ctypes.data)
    assert np.allclose(d.reshape(6, 6), _D(E, nu) / E, rtol=1e-13)
    O.lib().oracle_stiffness_3d_derivative(E, nu, 0.0, 1.0, d.ctypes.data)
    h = 1e-6
    fd = (_D(E, nu + h) - _D(E, nu - h)) / (2 * h)
    assert np.allclose(d.reshape(6, 6), fd, rtol=1e-7, atol=1e-3)
