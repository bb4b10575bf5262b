"""A long-lived GPU worker for the rule-compatible mode ("Threading").

The reference's rules start one `kmc` / `kmc_tools` process per rule instance (/root/reference/workflow/rules/
exp_type_1.smk:163,173,182,191,241,250,259 -- about 2 processes per (k, genome) plus a handful per (k, group)).  With the
drop-in shims every such process would create its own CUDA context (~0.3-0.5 s, far longer than the kernels of a 5 Mbp
genome), and concurrent Snakemake jobs would each hold one.  The worker owns ONE context; the shims, when the environment
variable KHB_WORKER_SOCKET names its UNIX socket, only forward their command line and working directory and relay the exit
status and stderr (this client side imports nothing but the standard library).  Requests are served one at a time -- calls
on a context are serialised by design (include/khoice_b200.h) -- so `snakemake --cores N` stays correct and simply queues.

    python -m khoice_b200.worker --socket /tmp/khb.sock [--device 0] &          # prints "khoice-b200 worker ready" when bound
    export KHB_WORKER_SOCKET=/tmp/khb.sock PATH=/path/to/khoice_b200/bin:$PATH
    snakemake --cores 8 ...                                                     # unmodified rules
    python -m khoice_b200.worker --socket /tmp/khb.sock --stop

Wire format: one JSON object per line each way: {"argv": [tool, ...], "cwd": path} -> {"rc": int, "stderr": text}.
There is no CPU fallback on either side: without a reachable worker the shim fails with exit status 1.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import socket
import sys
from typing import Callable, List, Optional

STOP = "__stop__"


def request(socket_path: str, argv: List[str], cwd: Optional[str] = None, timeout: Optional[float] = None) -> int:
    """Client: run `argv` (["kmc", ...] / ["kmc_tools", ...] / ["khb", ...]) in the worker; returns its exit status."""
    try:
        with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as s:
            s.settimeout(timeout)
            s.connect(socket_path)
            s.sendall((json.dumps({"argv": list(argv), "cwd": cwd or os.getcwd()}) + "\n").encode())
            buf = b""
            while not buf.endswith(b"\n"):
                chunk = s.recv(65536)
                if not chunk:
                    break
                buf += chunk
        reply = json.loads(buf.decode())
    except (OSError, ValueError) as e:
        print(f"{argv[0] if argv else 'khoice-b200'}: worker at {socket_path} is not reachable ({e}); there is no fallback", file=sys.stderr)
        return 1
    if reply.get("stderr"):
        sys.stderr.write(reply["stderr"])
    return int(reply.get("rc", 1))


def _reply(conn, rc: int, stderr: str) -> None:
    """A client that went away (a cancelled Snakemake job) must not take the worker down with it."""
    try:
        conn.sendall((json.dumps({"rc": rc, "stderr": stderr}) + "\n").encode())
    except OSError:
        pass


def serve(socket_path: str, handler: Callable[[List[str]], int], ready: Optional[Callable[[], None]] = None,
          healthy: Optional[Callable[[], bool]] = None, request_timeout: float = 30.0) -> int:
    """Server loop: accept, chdir to the client's working directory, run handler(argv), reply.  Returns the number of
    requests served when a client sends [STOP].  A client that disconnects or never finishes its request line (request_timeout
    seconds) is dropped; when `healthy()` turns false (a sticky CUDA error in the engine) the loop ends with SystemExit(3), so that
    the rules fail loudly instead of every later request failing against a dead context."""
    if os.path.exists(socket_path):
        os.remove(socket_path)
    served = 0
    home = os.getcwd()
    with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as srv:
        srv.bind(socket_path)
        srv.listen(64)
        if ready:
            ready()
        try:
            while True:
                conn, _ = srv.accept()
                with conn:
                    buf = b""
                    try:
                        conn.settimeout(request_timeout)
                        while not buf.endswith(b"\n"):
                            chunk = conn.recv(65536)
                            if not chunk:
                                break
                            buf += chunk
                        conn.settimeout(None)
                    except OSError:                                 # timeout / reset while reading the request: drop this client
                        continue
                    try:
                        req = json.loads(buf.decode())
                        argv = [str(a) for a in req["argv"]]
                    except (ValueError, KeyError, TypeError):
                        _reply(conn, 1, "khoice-b200 worker: malformed request\n")
                        continue
                    if argv == [STOP]:
                        _reply(conn, 0, "")
                        return served
                    err = io.StringIO()
                    rc = 1
                    try:
                        os.chdir(req.get("cwd") or home)           # the rules' paths are relative to Snakemake's workdir
                        with contextlib.redirect_stderr(err):
                            rc = int(handler(argv))
                    except BaseException as e:                      # the worker survives a failing request
                        err.write(f"khoice-b200 worker: {type(e).__name__}: {e}\n")
                        if isinstance(e, KeyboardInterrupt):
                            raise
                    finally:
                        os.chdir(home)
                    served += 1
                    _reply(conn, rc, err.getvalue())
                    if healthy is not None and not healthy():
                        raise SystemExit(3)
        finally:
            if os.path.exists(socket_path):
                os.remove(socket_path)


def main(argv: Optional[List[str]] = None) -> int:
    import argparse
    ap = argparse.ArgumentParser(description="long-lived B200 worker behind the kmc / kmc_tools shims")
    ap.add_argument("--socket", default=os.environ.get("KHB_WORKER_SOCKET"))
    ap.add_argument("--device", type=int, default=int(os.environ.get("KHB_DEVICE", "0")))
    ap.add_argument("--stop", action="store_true", help="ask the worker behind --socket to exit")
    ap.add_argument("--client", nargs=argparse.REMAINDER, help="forward this command line (used by khoice_b200/bin/*)")
    a = ap.parse_args(argv)
    if not a.socket:
        ap.error("--socket (or KHB_WORKER_SOCKET) is required")
    if a.stop:
        return request(a.socket, [STOP])
    if a.client is not None:
        return request(a.socket, a.client)
    from . import cli                      # the GPU side: only the serving process loads the engine
    from .engine import Engine
    os.environ.pop("KHB_WORKER_SOCKET", None)      # the handler must execute, not forward
    eng = Engine(a.device)                 # fails loudly without a B200
    cli.set_engine(eng)
    try:
        n = serve(a.socket, cli.main, ready=lambda: print("khoice-b200 worker ready", flush=True), healthy=eng.healthy)
        print(f"khoice-b200 worker: served {n} requests", flush=True)
    finally:
        cli.set_engine(None)
        eng.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
