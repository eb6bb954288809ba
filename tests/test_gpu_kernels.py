"""GPU: the two tcgen05 kernels against the SIMT checking kernels and fp64 torch, through the C ABI."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import kernel_check
    return kernel_check.load_lib()


def test_conv_and_wgrad_cases(lib):
    import kernel_check
    assert kernel_check.all_cases(lib)


def test_conv_large_tiles_many_waves(lib):
    """More tiles than SMs (persistent loop, both TMEM accumulator buffers, several phases)."""
    import kernel_check as kc
    assert kc.conv_case(lib, "5x5 s1 B16 Y40 X32 C128 N512", 16, 1, 40, 32, 128, 512, kc.taps_5x5_s1(), 16, 40, 32, 3, 128)
    assert kc.conv_case(lib, "same, bf16, blockN 256", 16, 1, 40, 32, 128, 512, kc.taps_5x5_s1(), 16, 40, 32, 1, 256)
    # split-K over many waves: 7 slices x (40 pair tiles / 160 single tiles), persistent loop
    assert kc.conv_case(lib, "5x5 s1 B16 Y40 X32 C128 N512", 16, 1, 40, 32, 128, 512, kc.taps_5x5_s1(), 16, 40, 32, 3, 256, ksplit=7)
    z0 = [(0, 0, 0, 0)]
    assert kc.wgrad_case(lib, "wgrad B16 Y40 X32 N512 C128", (16, 40, 32, 512), (16, 1, 40, 32, 128), kc.taps_5x5_s1(),
                         z0 * 25, 16, 40, 32, 3, 128, 4)
