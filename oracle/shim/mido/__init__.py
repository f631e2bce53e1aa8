"""mido stand-in: the three calls the reference makes (dxdata.py:315, 395-397).
TEST INFRASTRUCTURE ONLY."""


class Message:
    def __init__(self, type, data=()):
        assert type == "sysex"
        self.type = type
        self.data = tuple(int(b) for b in data)


def read_syx_file(path):
    with open(path, "rb") as f:
        raw = f.read()
    msgs = []
    i = 0
    while i < len(raw):
        if raw[i] == 0xF0:
            j = raw.index(0xF7, i)
            msgs.append(Message("sysex", raw[i + 1:j]))
            i = j + 1
        else:
            i += 1
    return msgs


def write_syx_file(path, messages):
    with open(path, "wb") as f:
        for m in messages:
            f.write(bytes([0xF0]) + bytes(m.data) + bytes([0xF7]))
