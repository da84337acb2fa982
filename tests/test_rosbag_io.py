"""ROS1 bag front end: container round trips, message codecs, event ordering.

No recorded bag ships with the reference (its ``testdata/`` is referenced but absent), so the fixtures are
written by ``BagWriter`` following the published v2.0 layout, and the structural invariants a ``rosbag``
reader relies on are asserted directly on the bytes."""
import struct

import numpy as np
import pytest

from pyratslam_b200 import rosbag_io as rb


def _run(T=12, hw=(16, 24), seed=0):
    rng = np.random.default_rng(seed)
    frames = rng.integers(0, 256, (T,) + hw, dtype=np.uint8)
    odom = np.stack([rng.uniform(0, 3, T), rng.uniform(-1, 1, T)], axis=1)
    return frames, odom


@pytest.mark.parametrize("compression", ["none", "bz2"])
@pytest.mark.parametrize("chunk_threshold", [1, 2000, 1 << 20])
def test_round_trip(tmp_path, compression, chunk_threshold):
    frames, odom = _run()
    path = str(tmp_path / "run.bag")
    rb.write_run(path, frames, odom, compression=compression, chunk_threshold=chunk_threshold)
    events = rb.read_events(path)
    assert [e[0] for e in events] == ["odom", "image"] * len(frames)
    f2, o2, stamps = rb.events_to_arrays(events)
    assert f2.dtype == np.uint8 and np.array_equal(f2, frames)
    assert np.array_equal(o2, odom)                      # float64 twists survive bit-exactly
    assert np.all(np.diff(stamps) > 0)


def test_container_layout(tmp_path):
    frames, odom = _run(T=5)
    path = str(tmp_path / "run.bag")
    rb.write_run(path, frames, odom, chunk_threshold=3000)
    buf = open(path, "rb").read()
    assert buf.startswith(b"#ROSBAG V2.0\n")
    recs = list(rb._records(buf, len(rb.MAGIC)))
    head, data, off = recs[0]
    assert head["op"] == bytes([rb.OP_BAG_HEADER]) and off == 13
    assert 13 + 4 + len(rb._pack_header(head)) + 4 + len(data) == 13 + 4096      # header record is 4096 bytes
    index_pos = struct.unpack("<Q", head["index_pos"])[0]
    n_conn = struct.unpack("<I", head["conn_count"])[0]
    n_chunk = struct.unpack("<I", head["chunk_count"])[0]
    ops = [r[0]["op"][0] for r in recs]
    assert ops.count(rb.OP_CHUNK) == n_chunk and n_chunk > 1
    tail = [r for r in recs if r[2] >= index_pos]
    assert [r[0]["op"][0] for r in tail] == [rb.OP_CONNECTION] * n_conn + [rb.OP_CHUNK_INFO] * n_chunk
    assert n_conn == 2
    # every chunk-info points at a chunk record, every index entry at a message record of its connection
    chunk_at = {r[2]: r for r in recs if r[0]["op"][0] == rb.OP_CHUNK}
    for h, _, _ in tail[n_conn:]:
        assert struct.unpack("<Q", h["chunk_pos"])[0] in chunk_at
    for i, (h, d, _) in enumerate(recs):
        if h["op"][0] != rb.OP_INDEX:
            continue
        j = i
        while recs[j][0]["op"][0] != rb.OP_CHUNK:
            j -= 1
        chunk = recs[j][1]
        conn = h["conn"]
        for e in range(struct.unpack("<I", h["count"])[0]):
            s, ns, moff = struct.unpack_from("<III", d, 12 * e)
            mh, _, _ = next(rb._records(chunk, moff))
            assert mh["op"] == bytes([rb.OP_MSG]) and mh["conn"] == conn and mh["time"] == struct.pack("<II", s, ns)
    bag = rb.BagReader(path)
    assert bag.topics == [rb.IMAGE_TOPIC, rb.ODOM_TOPIC] and len(bag) == 10 and bag.chunk_count == n_chunk
    assert {c["md5sum"] for c in bag.connections.values()} == {rb.MSG_TYPES["sensor_msgs/Image"][0],
                                                               rb.MSG_TYPES["nav_msgs/Odometry"][0]}


def test_time_order_and_topic_filter(tmp_path):
    path = str(tmp_path / "mixed.bag")
    with rb.BagWriter(path, chunk_threshold=200) as w:            # written out of order, across chunks
        w.write("/navbot/odom", "nav_msgs/Odometry", (5, 0), rb.encode_odometry(0.5, 0.0))
        w.write("other", "std_msgs/Int32", (1, 0), rb.encode_int32(-7))
        w.write("/navbot/odom", "nav_msgs/Odometry", (2, 500), rb.encode_odometry(0.25, -0.125))
        w.write("/navbot/camera/image", "sensor_msgs/Image", (2, 500), rb.encode_image(np.full((4, 4), 9, np.uint8)))
        w.write("navbot/experiencemap", "geometry_msgs/Pose2D", (3, 0), rb.encode_pose2d(1.5, -2.0, 0.25))
    bag = rb.BagReader(path)
    assert [m.stamp for m in bag.messages()] == [(1, 0), (2, 500), (2, 500), (3, 0), (5, 0)]
    assert rb.decode_int32(next(bag.messages(["other"])).data) == -7
    assert rb.decode_pose2d(next(bag.messages([rb.EM_TOPIC])).data) == (1.5, -2.0, 0.25)
    ev = rb.read_events(path)                                     # leading '/' ignored; ties keep file order
    assert [e[0] for e in ev] == ["odom", "image", "odom"]
    assert ev[0][2] == (0.25, -0.125) and ev[2][2] == (0.5, 0.0)
    f, o, _ = rb.events_to_arrays(ev)
    assert f.shape == (1, 4, 4) and np.array_equal(o, [[0.25, -0.125]])


def test_image_encodings():
    rng = np.random.default_rng(1)
    rgb = rng.integers(0, 256, (5, 7, 3), dtype=np.uint8)
    r, g, b = (rgb[..., i].astype(np.int64) for i in range(3))
    luma = ((r * 4899 + g * 9617 + b * 1868 + 8192) >> 14).astype(np.uint8)
    assert np.array_equal(rb.decode_image(rb.encode_image(rgb, encoding="rgb8")).image, luma)
    assert np.array_equal(rb.decode_image(rb.encode_image(rgb[..., ::-1], encoding="bgr8")).image, luma)
    assert np.abs(luma.astype(int) - np.rint(0.299 * r + 0.587 * g + 0.114 * b)).max() <= 1
    m16 = rng.integers(0, 65536, (3, 4), dtype=np.uint16)
    m16[0, :2] = (0, 65535)
    got = rb.decode_image(rb.encode_image(m16, encoding="mono16")).image
    assert np.array_equal(got, np.rint(m16 * (255.0 / 65535.0)).astype(np.uint8)) and got[0, 0] == 0 and got[0, 1] == 255
    # row padding (step > width) is skipped
    img = rng.integers(0, 256, (3, 5), dtype=np.uint8)
    msg = bytearray(rb.encode_image(img))
    padded = np.zeros((3, 8), np.uint8)
    padded[:, :5] = img
    head = rb._ser_header(0, (0, 0), "camera") + struct.pack("<II", 3, 5) + struct.pack("<I", 5) + b"mono8"
    raw = head + struct.pack("<BI", 0, 8) + struct.pack("<I", 24) + padded.tobytes()
    assert np.array_equal(rb.decode_image(raw).image, img) and bytes(msg[:len(head)]) == head
    with pytest.raises(rb.BagFormatError):
        rb.decode_image(head.replace(b"mono8", b"bayer") + struct.pack("<BI", 0, 8) + struct.pack("<I", 24) + padded.tobytes())


def test_odometry_codec():
    msg = rb.encode_odometry(1.25, -0.5, stamp=(3, 4), seq=9, position=(1, 2, 3))
    assert len(msg) == 12 + 4 + 4 + 4 + 9 + 56 + 288 + 48 + 288
    o = rb.decode_odometry(msg)
    assert o.stamp == (3, 4) and o.frame_id == "odom" and o.child_frame_id == "base_link"
    assert o.position == (1.0, 2.0, 3.0) and o.orientation == (0.0, 0.0, 0.0, 1.0)
    assert o.linear == (1.25, 0.0, 0.0) and o.angular == (0.0, 0.0, -0.5)


def test_malformed(tmp_path):
    p = tmp_path / "bad.bag"
    p.write_bytes(b"#ROSBAG V1.2\n")
    with pytest.raises(rb.BagFormatError):
        rb.BagReader(str(p))
    frames, odom = _run(T=3)
    good = tmp_path / "good.bag"
    rb.write_run(str(good), frames, odom)
    p.write_bytes(good.read_bytes()[:-5])
    with pytest.raises(rb.BagFormatError):
        rb.BagReader(str(p))
    with pytest.raises(ValueError):
        rb.BagWriter(str(p), compression="lz4")
