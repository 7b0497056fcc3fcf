"""Parser for what the reference receiver application prints (src/gmr1_rx.c), frame by frame (TEST INFRASTRUCTURE).

One dict per "[-]  FN:" line: fn, the control-channel burst of the frame (kind 'bcch' / 'ccch', crc, conv), and what
rx_tch3 (:538-600) did on the traffic channel: tch = 'dkab' | 'facch3' | 'tch3' | None, toa, bi / sync_id for FACCH3
bursts, `flush` = list of (crc, conv) printed by _rx_tch3_facch_flush (:394-452; two entries when the ciphered retry
ran), conv0 / conv1 / frame0 / frame1 for speech bursts, `assigned` = TN of an IMM.ASS seen in this frame, `end` = True
where the channel was released ("END @fn"); for rx_tch9 (:281-355): csd = 'tch9' | 'facch9', csd_toa, csd_sync, conv9 / avg
of a TCH9 block or csd_crc / csd_conv of a FACCH9 message."""
import re


def parse(lines):
    frames, cur, tag = [], None, None
    for l in lines:
        m = re.match(r"\[-\]  FN:\s*(-?\d+)", l)
        if m:
            cur = {"fn": int(m.group(1)), "kind": None, "crc": None, "conv": None, "tch": None, "flush": [],
                   "assigned": None, "end": False}
            frames.append(cur)
            tag = None
            continue
        if cur is None:
            continue
        if l.startswith("[.]   BCCH") or l.startswith("[.]   CCCH"):
            tag = cur["kind"] = l[6:10].lower()
        elif l.startswith("[.]   DKAB"):
            tag = cur["tch"] = "dkab"
        elif l.startswith("[.]   TCH9") or l.startswith("[.]   FACCH9"):
            tag = "csd"
            cur["csd"] = l[6:].strip().lower()
        elif l.startswith("[.]   FACCH3"):
            tag = cur["tch"] = "facch3"
            cur["bi"] = int(re.search(r"bi=(\d)", l).group(1))
        elif l.startswith("[.]   TCH3"):
            tag = cur["tch"] = "tch3"
        elif l.startswith("[+] TCH3 assigned on TN"):
            cur["assigned"] = int(l.split()[-1])
        elif l.startswith("END @"):
            cur["end"] = True
        elif l.startswith("toa="):
            m = re.match(r"toa=(-?[\d.]+)(?:, sync_id=(\d))?", l)
            if tag == "csd":
                cur["csd_toa"], cur["csd_sync"] = float(m.group(1)), int(m.group(2))
                continue
            cur["toa"] = float(m.group(1))
            if m.group(2) is not None:
                cur["sync_id"] = int(m.group(2))
        elif l.startswith("crc="):
            m = re.match(r"crc=(-?\d+), conv=(-?\d+)", l)
            if tag == "csd":
                cur["csd_crc"], cur["csd_conv"] = int(m.group(1)), int(m.group(2))
            elif tag == "facch3":
                cur["flush"].append((int(m.group(1)), int(m.group(2))))
            else:
                cur["crc"], cur["conv"] = int(m.group(1)), int(m.group(2))
        elif l.startswith("fn=") and "conv9=" in l:
            m = re.match(r"fn=(-?\d+), conv9=(-?\d+), avg=(-?\d+)", l)
            cur["conv9"], cur["avg"] = int(m.group(2)), int(m.group(3))
        elif l.startswith("conv="):
            m = re.match(r"conv=\s*(-?\d+),\s*(-?\d+)", l)
            cur["conv0"], cur["conv1"] = int(m.group(1)), int(m.group(2))
        elif l.startswith("frame0="):
            cur["frame0"] = bytes.fromhex(l[7:].strip())
        elif l.startswith("frame1="):
            cur["frame1"] = bytes.fromhex(l[7:].strip())
    return frames
