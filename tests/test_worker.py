"""The long-lived worker behind the kmc / kmc_tools shims (khoice_b200/worker.py), host side only: the wire protocol, the
working-directory hand-over, exit status and stderr relay, survival of a failing request, the shell shims' forwarding.
The GPU side (a real Engine behind the socket) is covered by tests/test_gpu_workflow.py."""
import os
import subprocess
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _start(sock, handler):
    from khoice_b200 import worker
    ready = threading.Event()
    out = {}
    t = threading.Thread(target=lambda: out.setdefault("served", worker.serve(sock, handler, ready=ready.set)), daemon=True)
    t.start()
    assert ready.wait(10)
    return t, out


def test_requests_run_in_the_clients_directory_and_relay_status_and_stderr(tmp_path):
    from khoice_b200 import worker
    seen = []

    def handler(argv):
        seen.append((list(argv), os.getcwd()))
        if argv[0] == "boom":
            raise RuntimeError("kernel exploded")
        if "bad" in argv:
            print("kmc_tools (khoice-b200): no such database", file=sys.stderr)
            return 1
        open("made_here.txt", "w").write("x")          # relative path: resolved in the CLIENT's directory
        return 0

    sock = str(tmp_path / "w.sock")
    t, out = _start(sock, handler)
    work = tmp_path / "workdir"
    work.mkdir()
    assert worker.request(sock, ["kmc", "-fm", "-k31", "-ci1", "a.fna.gz", "out", "tmp/"], cwd=str(work)) == 0
    assert (work / "made_here.txt").exists() and seen[0] == (["kmc", "-fm", "-k31", "-ci1", "a.fna.gz", "out", "tmp/"], str(work))
    assert worker.request(sock, ["kmc_tools", "transform", "bad", "histogram", "h.txt"], cwd=str(work)) == 1
    assert worker.request(sock, ["boom"], cwd=str(work)) == 1           # the worker survives an exception in a request
    assert worker.request(sock, ["kmc", "again"], cwd=str(work)) == 0
    assert worker.request(sock, [worker.STOP]) == 0
    t.join(10)
    assert out["served"] == 4 and not os.path.exists(sock)
    assert worker.request(sock, ["kmc", "x"]) == 1                      # nobody listens any more: hard failure, no fallback


def test_shell_shims_forward_when_the_socket_variable_is_set(tmp_path):
    """khoice_b200/bin/kmc_tools with KHB_WORKER_SOCKET: the command line (options with dashes included), the caller's
    directory and the exit status travel; the client process never loads the engine."""
    got = []

    def handler(argv):
        got.append((list(argv), os.getcwd()))
        print("from the worker", file=sys.stderr)
        return 7 if argv[-1] == "fail" else 0

    sock = str(tmp_path / "s.sock")
    t, out = _start(sock, handler)
    env = dict(os.environ, KHB_WORKER_SOCKET=sock)
    shim = os.path.join(ROOT, "khoice_b200", "bin", "kmc_tools")
    r = subprocess.run([shim, "simple", "A", "B", "intersect", "O", "-ocsum"], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "from the worker" in r.stderr, r.stderr
    assert got[0] == (["kmc_tools", "simple", "A", "B", "intersect", "O", "-ocsum"], str(tmp_path))
    r = subprocess.run([os.path.join(ROOT, "khoice_b200", "bin", "kmc"), "-fm", "-k21", "fail"], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=60)
    assert r.returncode == 7 and got[1][0] == ["kmc", "-fm", "-k21", "fail"]
    r = subprocess.run([sys.executable, "-m", "khoice_b200.worker", "--socket", sock, "--stop"], cwd=ROOT, capture_output=True, text=True, timeout=60)
    assert r.returncode == 0
    t.join(10)
    assert out["served"] == 2


def test_worker_survives_clients_that_vanish_or_stall_and_stops_when_unhealthy(tmp_path):
    """A client killed in mid-request (a cancelled Snakemake job), one that connects and never sends its line, and one that is gone when
    the reply is due must not take the shared worker down; an engine with a sticky error ends the loop (SystemExit 3)."""
    import socket
    import time
    from khoice_b200 import worker
    state = {"ok": True}

    def handler(argv):
        if argv[0] == "slow":
            time.sleep(0.3)
        if argv[0] == "poison":
            state["ok"] = False
        return 0

    sock = str(tmp_path / "w.sock")
    ready = threading.Event()
    out = {}

    def run():
        try:
            out["served"] = worker.serve(sock, handler, ready=ready.set, healthy=lambda: state["ok"], request_timeout=0.5)
        except SystemExit as e:
            out["exit"] = e.code

    t = threading.Thread(target=run, daemon=True)
    t.start()
    assert ready.wait(10)
    c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)            # half a request, then gone
    c.connect(sock)
    c.sendall(b'{"argv": ["km')
    c.close()
    c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)            # connects and never sends a newline: dropped after the timeout
    c.connect(sock)
    c.sendall(b'{"argv": ["kmc"')
    assert worker.request(sock, ["kmc", "still", "alive"]) == 0
    c.close()
    c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)            # gone before the reply is written
    c.connect(sock)
    c.sendall(b'{"argv": ["slow"], "cwd": null}\n')
    c.close()
    assert worker.request(sock, ["kmc", "after", "the", "broken", "pipe"]) == 0
    assert worker.request(sock, ["poison"]) == 0                      # the request itself is answered, then the worker leaves
    t.join(10)
    assert out.get("exit") == 3
