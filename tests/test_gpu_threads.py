"""Re-entrancy of the C ABI (horde-ad.cabal:126-145: `-threaded`, tests with -N8): eight host
threads hammer two-kernel entry points that share the reduction / scatter scratch (`hb_sum_all`, `hb_dot_all`,
`hb_scatter_add`, the gather generator) and release their handles concurrently; every result must equal the serial
one bit for bit.  ctypes drops the GIL around every foreign call, so the threads really do interleave in the library."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_eight_host_threads_hammer_the_entry_points(dev, oracle):
    rng = np.random.default_rng(77)
    n_threads, rounds = 8, 12
    xs = [rng.uniform(-1, 1, (300_000 + 1000 * t,)) for t in range(n_threads)]
    ys = [rng.uniform(-.5, .5, (48, 3, 64 + t)) for t in range(n_threads)]
    f = lambda i: [i[0] + i[1]]
    # serial results first
    serial = []
    for x, y in zip(xs, ys):
        dx = dev.sconcrete(x)
        serial.append((dev.to_numpy(dev.ssum0(dx)), dev.to_numpy(dev.sdot0(dx, dx)),
                       dev.to_numpy(dev.sscatter([50], dev.sconcrete(y), f, 2)),
                       dev.to_numpy(dev.sgather([40, 3], dev.sconcrete(y), f, 1))))
    errors = []
    start = threading.Barrier(n_threads)

    def work(t):
        try:
            start.wait()
            for _ in range(rounds):
                dx, dy = dev.sconcrete(xs[t]), dev.sconcrete(ys[t])
                got = (dev.to_numpy(dev.ssum0(dx)), dev.to_numpy(dev.sdot0(dx, dx)),
                       dev.to_numpy(dev.sscatter([50], dy, f, 2)), dev.to_numpy(dev.sgather([40, 3], dy, f, 1)))
                for g, s in zip(got, serial[t]):
                    if not np.array_equal(g, s):
                        errors.append((t, "result differs from the serial run"))
                del dx, dy          # hb_release from this thread
        except Exception as e:      # noqa: BLE001
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=300)
    assert not errors, errors[:3]
    np.testing.assert_array_equal(serial[0][2], oracle.sscatter([50], ys[0], f, 2))


def test_graph_pool_blocks_return_to_the_general_pool(dev):
    """hb_graph_free with arrays of the capture still alive: their blocks must rejoin the general pool when they
    are released later, i.e. a same-sized eager allocation afterwards does not grow the reservation."""
    import ctypes as C
    from horde_ad_b200 import ffi
    x = dev.sconcrete(np.ones(1 << 20))
    _warm = dev.binary("add", x, x)
    dev.sync()
    ffi.call("hb_graph_begin")
    try:
        y = dev.binary("add", x, x)
    finally:
        g = C.c_void_p()
        ffi.call("hb_graph_end", C.byref(g))
    ffi.call("hb_graph_launch", g)
    dev.sync()
    ffi.call("hb_graph_free", g)
    live, res0 = C.c_uint64(), C.c_uint64()
    del y, _warm                       # released AFTER the graph is gone
    ffi.call("hb_mem_stats", C.byref(live), C.byref(res0))
    z1, z2 = dev.binary("add", x, x), dev.binary("mul", x, x)
    res1 = C.c_uint64()
    ffi.call("hb_mem_stats", C.byref(live), C.byref(res1))
    assert res1.value == res0.value, (res0.value, res1.value)
    assert dev.to_numpy(z1)[0] == 2.0 and dev.to_numpy(z2)[0] == 1.0
