"""Remote-viewer endpoint with the module-level interface of the reference's ``gaussian_renderer/network_gui.py``
(used by train_gui.py:216-229): ``init(host, port)``, ``try_connect()``, ``receive()``, ``send(image_bytes, verify)``
and the module attributes ``conn`` / ``addr`` / ``host`` / ``port`` that the trainer reads and resets.

Wire format (SIBR remote viewer): every message is a 4-byte little-endian length followed by that many bytes.  The
viewer sends a JSON object (resolution, field of view, clip planes, the two 4x4 matrices flattened row-major, and a few
flags); the trainer answers with the raw image bytes (if any) and a length-prefixed ASCII verification string.
Nothing here touches the render hot path; the module exists so that the package is a complete stand-in."""
import json
import socket
import traceback

import torch

host = "127.0.0.1"
port = 6009
conn = None
addr = None
listener = None


def _listener():
    global listener
    if listener is None:
        listener = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
    return listener


def init(wish_host, wish_port):
    """Bind and listen without blocking the training loop (accept() is polled by try_connect)."""
    global host, port
    host, port = wish_host, wish_port
    s = _listener()
    s.bind((host, port))
    s.listen()
    s.settimeout(0)


def try_connect():
    global conn, addr
    try:
        conn, addr = _listener().accept()
        print(f"\nConnected by {addr}")
        conn.settimeout(None)
    except Exception:
        pass       # nobody is knocking: the trainer polls again next iteration


def _recv_exact(n):
    chunks, got = [], 0
    while got < n:
        c = conn.recv(n - got)
        if not c:
            raise ConnectionError("viewer closed the connection")
        chunks.append(c)
        got += len(c)
    return b"".join(chunks)


def read():
    n = int.from_bytes(_recv_exact(4), "little")
    return json.loads(_recv_exact(n).decode("utf-8"))


def send(message_bytes, verify):
    if message_bytes is not None:
        conn.sendall(message_bytes)
    conn.sendall(len(verify).to_bytes(4, "little"))
    conn.sendall(bytes(verify, "ascii"))


def receive():
    """-> (MiniCam, do_training, do_shs_python, do_rot_scale_python, keep_alive, scaling_modifier), all None for an
    empty-resolution message.  The viewer's camera looks down -z with y up; the y and z columns are flipped to the
    renderer's convention (as the reference does)."""
    msg = read()
    width, height = msg["resolution_x"], msg["resolution_y"]
    if width == 0 or height == 0:
        return None, None, None, None, None, None
    try:
        from scene.cameras import MiniCam          # the reference's camera container (scene/cameras.py)
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        wv = torch.tensor(msg["view_matrix"]).reshape(4, 4).to(dev)
        wv[:, 1] = -wv[:, 1]
        wv[:, 2] = -wv[:, 2]
        full = torch.tensor(msg["view_projection_matrix"]).reshape(4, 4).to(dev)
        full[:, 1] = -full[:, 1]
        cam = MiniCam(width, height, msg["fov_y"], msg["fov_x"], msg["z_near"], msg["z_far"], wv, full)
        return (cam, bool(msg["train"]), bool(msg["shs_python"]), bool(msg["rot_scale_python"]), bool(msg["keep_alive"]),
                msg["scaling_modifier"])
    except Exception:
        print("")
        traceback.print_exc()
        raise
