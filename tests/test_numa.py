"""Host placement helper (afivo_streamer_b200/numa.py) against a fake sysfs tree: pure host logic."""
import os

from afivo_streamer_b200 import numa


def test_parse_cpulist():
    assert numa.parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert numa.parse_cpulist("") == set()
    assert numa.parse_cpulist("5") == {5}


def test_node_lookup_and_binding_on_a_fake_sysfs(tmp_path, monkeypatch):
    addr = "0000:1b:00.0"
    d = tmp_path / "bus/pci/devices" / addr
    d.mkdir(parents=True)
    (d / "numa_node").write_text("1\n")
    n = tmp_path / "devices/system/node/node1"
    n.mkdir(parents=True)
    allowed = sorted(os.sched_getaffinity(0))
    (n / "cpulist").write_text(f"{allowed[0]}\n")
    monkeypatch.setattr(numa, "pci_address", lambda i: addr)
    assert numa.gpu_numa_node(0, str(tmp_path)) == 1
    assert numa.node_cpus(1, str(tmp_path)) == {allowed[0]}
    before = os.sched_getaffinity(0)
    try:
        assert numa.bind_to_gpu_node(0, str(tmp_path)) == 1
        assert os.sched_getaffinity(0) == {allowed[0]}
    finally:
        os.sched_setaffinity(0, before)
    # a node without any allowed core, an unknown node and a single-node machine leave the thread alone
    (n / "cpulist").write_text("100000\n")
    assert numa.bind_to_gpu_node(0, str(tmp_path)) is None
    (d / "numa_node").write_text("-1\n")
    assert numa.gpu_numa_node(0, str(tmp_path)) is None
    assert numa.bind_to_gpu_node(0, str(tmp_path)) is None
    assert os.sched_getaffinity(0) == before
