import sys, json, random, struct
sys.path.insert(0, '/root/repo')
from tests.test_native_ingest import rand_doc
seed = int(sys.argv[1]); rng = random.Random(seed)
def style(d):
    t, kw = rng.random(), {"ensure_ascii": rng.random() < 0.5}
    if t < 0.35: kw["separators"] = (",", ":")
    elif t < 0.55: kw["indent"] = rng.choice([1, 2, "\t"])
    s = json.dumps(d, **kw)
    r = rng.random()
    if r < 0.05: s = s[:rng.randint(0, len(s))]                      # truncated
    elif r < 0.10:                                                    # byte noise
        b = bytearray(s.encode()); 
        for _ in range(rng.randint(1, 4)):
            if b: b[rng.randrange(len(b))] = rng.choice(b'{}[]",:\\ 0e-.x\n')
        return bytes(b)
    return s.encode()
cells = [style(rand_doc(rng)) for _ in range(400)] + [b"", b"{", b"[]", b"null", b'{"objects": [{"name": "a"}]}', b'{"objects":[{"name":"\\u4e2d","polygon":{"ptList":[{"x":1,"y":2}]}}]}']
with open(sys.argv[2], "wb") as f:
    f.write(struct.pack("<Q", len(cells)))
    for c in cells: f.write(struct.pack("<Q", len(c))); f.write(c)
atoms = [b",", b",", b'"', b'""', b"\n", b"\n", b"\r\n", b" ", b"a", b"http://x", b'{"k": 1}', b"1", b"NA", b"\xe4\xb8\xad", b"\xff", b"\r"]
rows = []
for _ in range(200):
    rows.append(b"".join(rng.choice(atoms) for _ in range(rng.randint(1, 30))))
open(sys.argv[3], "wb").write(b"a,b,c\n" + b"\n".join(rows))
