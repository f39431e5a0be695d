"""TEST INFRASTRUCTURE: a reader for the subset of Fressian that maelstrom.net.journal's handlers
produce (src/maelstrom/net/journal.clj:55-113): packed ints, packed-length strings, the "key" tag,
STRUCTTYPE / struct cache, PUT_PRIORITY_CACHE / priority cache, closed lists, the "map" tag.
Written from the published org.fressian encoding (the library is a third-party dependency that is not
under /root/reference and there is no JVM here), independently of csrc/ms_fressian.h's writer."""
import io


class Keyword(str):
    pass


class FressianReader:
    def __init__(self, data):
        self.f = io.BytesIO(data)
        self.pcache = []
        self.structs = []

    def _u8(self):
        b = self.f.read(1)
        if not b:
            raise EOFError
        return b[0]

    def _raw(self, n):
        return int.from_bytes(self.f.read(n), "big")

    def read_int(self):
        v = self.read_object()
        assert isinstance(v, int)
        return v

    def _int_from_code(self, c):
        if c == 0xFF:
            return -1
        if c <= 0x3F:
            return c
        if 0x40 <= c <= 0x5F:
            return ((c - 0x50) << 8) | self._raw(1)
        if 0x60 <= c <= 0x6F:
            return ((c - 0x68) << 16) | self._raw(2)
        if 0x70 <= c <= 0x73:
            return ((c - 0x72) << 24) | self._raw(3)
        if 0x74 <= c <= 0x77:
            return ((c - 0x76) << 32) | self._raw(4)
        if 0x78 <= c <= 0x7B:
            return ((c - 0x7A) << 40) | self._raw(5)
        if 0x7C <= c <= 0x7F:
            return ((c - 0x7E) << 48) | self._raw(6)
        if c == 0xF8:
            v = self._raw(8)
            return v - (1 << 64) if v >> 63 else v
        return None

    def _struct(self, tag, n):
        if tag == "ev":
            assert n == 4
            return {"id": self.read_int(), "time": self.read_int(), "type": self.read_object(), "message": self.read_object()}
        if tag == "msg":
            assert n == 4
            return {"id": self.read_int(), "src": self.read_object(), "dest": self.read_object(), "body": self.read_object()}
        raise ValueError("unknown struct tag %r" % tag)

    def read_object(self):
        c = self._u8()
        v = self._int_from_code(c)
        if v is not None:
            return v
        if 0x80 <= c <= 0x9F:
            return self.pcache[c - 0x80]
        if 0xA0 <= c <= 0xAF:
            tag, n = self.structs[c - 0xA0]
            return self._struct(tag, n)
        if c == 0xCC:
            return self.pcache[self.read_int()]
        if c == 0xCD:                                      # PUT_PRIORITY_CACHE: the slot is reserved first
            idx = len(self.pcache)
            self.pcache.append(None)
            self.pcache[idx] = self.read_object()
            return self.pcache[idx]
        if c == 0xCA:                                      # "key": namespace, name
            ns, name = self.read_object(), self.read_object()
            assert ns is None
            return Keyword(name)
        if c == 0xF7:
            return None
        if 0xDA <= c <= 0xE1:
            return self.f.read(c - 0xDA).decode("utf-8")
        if c == 0xE3:
            return self.f.read(self.read_int()).decode("utf-8")
        if c == 0xEF:                                      # STRUCTTYPE tag n, then the components
            tag, n = self.read_object(), self.read_int()
            self.structs.append((tag, n))
            return self._struct(tag, n)
        if c == 0xF0:
            tag, n = self.structs[self.read_int()]
            return self._struct(tag, n)
        if c == 0xC0:                                      # "map": one component, a list of k v k v ...
            items = self.read_object()
            return dict(zip(items[0::2], items[1::2]))
        if c == 0xED:                                      # closed list ... END_COLLECTION
            out = []
            while True:
                p = self.f.tell()
                if self._u8() == 0xFD:
                    return out
                self.f.seek(p)
                out.append(self.read_object())
        if 0xE4 <= c <= 0xEB:
            return [self.read_object() for _ in range(c - 0xE4)]
        raise ValueError("unhandled Fressian code 0x%02x at %d" % (c, self.f.tell() - 1))

    def read_all(self):
        out = []
        while True:
            try:
                out.append(self.read_object())
            except EOFError:
                return out
