"""The oracle's dilated causal conv against the reference's own known answers
(the deterministic prints of ops.py:243-254; values listed in SURVEY.md 8(c))."""
import numpy as np

from oracle import srwn_oracle as orc


def _x(kat):
    return kat["x"].astype(np.float64).reshape(1, -1, 1)


def test_kat_filters_dilation1(conv_kat):
    x = _x(conv_kat)
    for key, f in (("l243", [1, 1]), ("l244", [1, 0, 1]), ("l245", [1, 0, 0, 0, 1])):
        w = np.array(f, np.float64).reshape(len(f), 1, 1)
        out = orc.dilated_causal_conv1d(x, w)
        np.testing.assert_array_equal(out.reshape(-1), conv_kat[key])


def test_kat_dilations(conv_kat):
    x = _x(conv_kat)
    w = np.ones((2, 1, 1))
    for key, d in (("l246", 2), ("l247", 3), ("l248", 4), ("l249", 6)):
        out = orc.dilated_causal_conv1d(x, w, dilation_rate=d)
        np.testing.assert_array_equal(out.reshape(-1), conv_kat[key])


def test_kat_two_output_channels(conv_kat):
    x = _x(conv_kat)
    f4 = np.array([[1, 2, 1, 2]], np.float64).reshape(2, 1, 2)       # ops.py:229,252
    np.testing.assert_array_equal(orc.dilated_causal_conv1d(x, f4)[0], conv_kat["l252"])
    np.testing.assert_array_equal(orc.valid_conv1d(x, f4)[0], conv_kat["l254"])  # ops.py:254


def test_kat_equivalences():
    # f=[1,0,1] (K=3,d=1) == f=[1,1] with d=2; f=[1,0,0,0,1] == d=4 (ops.py:244/246, 245/248)
    x = np.arange(1, 9, dtype=np.float64).reshape(1, -1, 1)
    a = orc.dilated_causal_conv1d(x, np.array([1, 0, 1.]).reshape(3, 1, 1))
    b = orc.dilated_causal_conv1d(x, np.ones((2, 1, 1)), 2)
    np.testing.assert_array_equal(a, b)
