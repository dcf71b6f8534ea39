"""Material::scatter / emit: device vs oracle per material (tests/scatter_cases.py), on the host build of the device
header here and through rt_scatter_batch on the GPU."""
import ctypes as C

import numpy as np
import pytest

import mu_lambda_raytracer_b200 as rt
from mu_lambda_raytracer_b200 import abi
import scatter_cases as SC
import support as S


def test_scatter_matches_oracle_emulated():
    def run(desc, arr):
        es = S.EmulScene(desc)
        out = (abi.RtScatterOut * len(arr))()
        S.emul().emul_scatter_batch(es.h, arr, len(arr), out)
        return np.ctypeslib.as_array(out).copy()
    assert SC.check_all(run, n=60_000)


@pytest.mark.gpu
def test_scatter_matches_oracle_on_device():
    def run(desc, arr):
        scene = rt.Scene(rt.SceneDescription(desc, owned=False))
        out = (abi.RtScatterOut * len(arr))()
        abi.check(abi.load().rt_scatter_batch(scene.handle, arr, len(arr), out))
        scene.close()
        return np.ctypeslib.as_array(out).copy()
    assert SC.check_all(run, n=1_000_000)
