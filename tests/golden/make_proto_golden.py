#!/usr/bin/env python
"""Mechanical conformance fixture for lzy_b200/grpc_api/protos.py (missing #8): parse the reference's own
.proto files — the slots API, both channel-manager services and everything they import that the kept wire contract
uses — into a plain JSON description (packages, messages with field name / number / type / label / oneof, nested types,
enums, services with their rpcs and streaming flags).  tests/test_proto_conformance.py holds protos.py to it.

Run in the authoring container (needs /root/reference):  python tests/golden/make_proto_golden.py
The parser is a small tokenizer + recursive descent over the proto3 subset these files use; it never imports or
executes anything from the reference."""
import json
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
REF = Path("/root/reference")
FILES = [
    "lzy/slots-api/src/main/proto/ai/lzy/v1/slots/slots-api.proto",
    "lzy/channel-manager-api/src/main/proto/ai/lzy/v1/channel/channel-manager.proto",
    "lzy/channel-manager-api/src/main/proto/ai/lzy/v1/channel/channel-manager-private.proto",
    "lzy/worker-api/src/main/proto/ai/lzy/v1/worker/worker-service.proto",
    "long-running-api/src/main/proto/ai/lzy/v1/long-running/operation.proto",
    "model/src/main/proto/ai/lzy/v1/common/channel.proto",
    "model/src/main/proto/ai/lzy/v1/common/data-scheme.proto",
    "model/src/main/proto/ai/lzy/v1/common/storage.proto",
    "model/src/main/proto/ai/lzy/v1/common/slot.proto",
    "model/src/main/proto/ai/lzy/v1/common/env.proto",
    "model/src/main/proto/ai/lzy/v1/common/operation.proto",
]
TOKEN = re.compile(r'"(?:[^"\\]|\\.)*"|[A-Za-z_][\w.]*|-?\d+|[{}()\[\]=;,<>]')


def tokens(text: str):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    return TOKEN.findall(text)


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else None

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def expect(self, tok):
        got = self.next()
        assert got == tok, f"expected {tok!r}, got {got!r} at token {self.i}"

    def skip_statement(self):
        depth = 0
        while True:
            tok = self.next()
            if tok == "{":
                depth += 1
            elif tok == "}":
                depth -= 1
                if depth == 0:
                    return
            elif tok == ";" and depth == 0:
                return

    def skip_options(self):
        if self.peek() == "[":
            while self.next() != "]":
                pass

    def file(self):
        out = {"package": "", "imports": [], "messages": {}, "enums": {}, "services": {}}
        while self.peek() is not None:
            tok = self.next()
            if tok == "package":
                out["package"] = self.next()
                self.expect(";")
            elif tok == "import":
                name = self.next()
                if name in ("public", "weak"):
                    name = self.next()
                out["imports"].append(name.strip('"'))
                self.expect(";")
            elif tok in ("syntax", "option"):
                self.skip_statement()
            elif tok == "message":
                self.message(out, "")
            elif tok == "enum":
                self.enum(out, "")
            elif tok == "service":
                self.service(out)
            else:
                raise AssertionError(f"unexpected top-level token {tok!r}")
        return out

    def message(self, out, scope):
        name = self.next()
        full = f"{scope}{name}"
        fields = []
        out["messages"][full] = fields
        self.expect("{")
        self.body(out, full, fields, None)

    def body(self, out, full, fields, oneof):
        while True:
            tok = self.next()
            if tok == "}":
                return
            if tok == "message":
                self.message(out, full + ".")
            elif tok == "enum":
                self.enum(out, full + ".")
            elif tok == "oneof":
                oname = self.next()
                self.expect("{")
                self.body(out, full, fields, oname)
            elif tok in ("option", "reserved", "extensions"):
                self.i -= 1
                self.skip_statement()
            elif tok == "map":
                self.expect("<")
                k = self.next()
                self.expect(",")
                v = self.next()
                self.expect(">")
                fname = self.next()
                self.expect("=")
                num = int(self.next())
                self.skip_options()
                self.expect(";")
                fields.append({"name": fname, "number": num, "type": f"map<{k},{v}>", "label": "repeated", "oneof": oneof})
            else:
                label = "optional"
                if tok in ("repeated", "optional", "required"):
                    label = "repeated" if tok == "repeated" else "optional"
                    tok = self.next()
                ftype = tok
                fname = self.next()
                self.expect("=")
                num = int(self.next())
                self.skip_options()
                self.expect(";")
                fields.append({"name": fname, "number": num, "type": ftype, "label": label, "oneof": oneof})

    def enum(self, out, scope):
        name = self.next()
        values = {}
        out["enums"][f"{scope}{name}"] = values
        self.expect("{")
        while True:
            tok = self.next()
            if tok == "}":
                return
            if tok in ("option", "reserved"):
                self.i -= 1
                self.skip_statement()
                continue
            self.expect("=")
            values[tok] = int(self.next())
            self.skip_options()
            self.expect(";")

    def service(self, out):
        name = self.next()
        rpcs = {}
        out["services"][name] = rpcs
        self.expect("{")
        while True:
            tok = self.next()
            if tok == "}":
                return
            if tok == ";":  # `rpc X (A) returns (B) {};`
                continue
            if tok == "option":
                self.i -= 1
                self.skip_statement()
                continue
            assert tok == "rpc", tok
            rname = self.next()
            self.expect("(")
            cstream = self.peek() == "stream"
            if cstream:
                self.next()
            req = self.next()
            self.expect(")")
            assert self.next() == "returns"
            self.expect("(")
            sstream = self.peek() == "stream"
            if sstream:
                self.next()
            resp = self.next()
            self.expect(")")
            if self.peek() == "{":
                self.skip_statement()
            else:
                self.expect(";")
            rpcs[rname] = {"request": req, "response": resp, "client_streaming": cstream, "server_streaming": sstream}


def main():
    out = {}
    for rel in FILES:
        text = (REF / rel).read_text()
        out[rel] = Parser(tokens(text)).file()
    dest = ROOT / "tests/golden/reference_proto_vectors.json"
    dest.write_text(json.dumps(out, indent=1, sort_keys=True))
    n_msgs = sum(len(f["messages"]) for f in out.values())
    n_rpcs = sum(len(r) for f in out.values() for r in f["services"].values())
    print(f"{dest.name}: {len(out)} files, {n_msgs} messages, {n_rpcs} rpcs")


if __name__ == "__main__":
    main()
