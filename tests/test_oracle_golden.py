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


def test_plate_section_matrices_golden(built):
    # /root/reference/tests/property/mast_solid_2d_section_element_property_card.cpp:614-618, 660-664 (extension),
    # :706-710, 752-756 (bending), :799-803, 846-850 (extension-bending), :893-894 (transverse shear)
    # and the material C of tests/material/mast_isotropic_material.cpp:345-350 (plane stress), :400-405 (plane strain)
    # for E=72e9, nu=0.33, h=0.06, off=0.03, kappa=5/6
    def mats(ps):
        A, B, D, S = np.zeros(9), np.zeros(9), np.zeros(9), np.zeros(4)
        O.lib().oracle_plate_section_matrices(72e9, 0.33, 0.06, 0.03, 5.0 / 6.0, ps, A.ctypes.data, B.ctypes.data,
                                              D.ctypes.data, S.ctypes.data)
        return A.reshape(3, 3), B.reshape(3, 3), D.reshape(3, 3), S.reshape(2, 2)

    def sym(d, o, s):
        return np.array([[d, o, 0.0], [o, d, 0.0], [0.0, 0.0, s]])

    A, B, D, S = mats(0)
    assert np.allclose(A, sym(4.847940747390865e09, 1.599820446638986e09, 1.624060150375940e09), rtol=1e-14, atol=0)
    assert np.allclose(D, sym(5.817528896869037e06, 1.919784535966782e06, 1.948872180451127e06), rtol=1e-14, atol=0)
    assert np.allclose(B, sym(1.454382224217259e08, 0.479946133991696e08, 0.487218045112782e08), rtol=1e-14, atol=0)
    assert np.allclose(S, np.eye(2) * 1.353383458646616e09, rtol=1e-14, atol=0)
    assert np.allclose(A / 0.06, sym(8.079901245651442e10, 2.666367411064976e10, 2.706766917293233e10), rtol=1e-14, atol=0)
    A, B, D, S = mats(1)
    assert np.allclose(A, sym(6.400707651481644e09, 3.152587350729766e09, 1.624060150375940e09), rtol=1e-14, atol=0)
    assert np.allclose(D, sym(7.680849181777972e06, 3.783104820875719e06, 1.948872180451128e06), rtol=1e-14, atol=0)
    assert np.allclose(B, sym(1.920212295444493e08, 0.945776205218930e08, 0.487218045112782e08), rtol=1e-14, atol=0)
