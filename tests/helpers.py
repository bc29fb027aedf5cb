"""Shared test helpers: an independent decode of response bytes (dnspython) into the
semantic tuples oracle/binder_ref.py produces, and a minimal query-packet splitter."""
import json
import struct

import dns.flags
import dns.message
import dns.rdatatype

import binder_ref as R


def snapshot(entries):
    """[(path, data)] -> JSONL bytes."""
    return ('\n'.join(json.dumps({'path': p, 'data': d}) for p, d in entries) + '\n').encode()


def split_query(pkt):
    """(labels, qtype, rd, opcode, edns) of a well-formed query packet."""
    flags = struct.unpack('>H', pkt[2:4])[0]
    pos, labels = 12, []
    while pkt[pos]:
        l = pkt[pos]
        labels.append(bytes(pkt[pos + 1:pos + 1 + l]))
        pos += 1 + l
    qtype = struct.unpack('>H', pkt[pos + 1:pos + 3])[0]
    arcount = struct.unpack('>H', pkt[10:12])[0]
    return labels, qtype, bool(flags & 0x0100), (flags >> 11) & 0xF, arcount > 0


def _name(n):
    """dns.name.Name -> latin-1 view of the raw labels joined by '.' (no escaping)."""
    return '.'.join(l.decode('latin-1') for l in n.labels if l)


def decode_semantic(wire):
    """Response bytes -> (rcode, answers, authority, additional, flags dict) with RR tuples
    shaped like binder_ref.Response.as_tuple()."""
    m = dns.message.from_wire(wire, one_rr_per_rrset=True, raise_on_truncation=False)
    def rrs(section):
        out = []
        for rrset in section:
            owner = _name(rrset.name)
            for rd in rrset:
                if rrset.rdtype == dns.rdatatype.A:
                    out.append((owner, rrset.ttl, 'A', rd.address))
                elif rrset.rdtype == dns.rdatatype.SRV:
                    assert rd.priority == 0 and rd.weight == 10
                    out.append((owner, rrset.ttl, 'SRV', rd.port, _name(rd.target)))
                elif rrset.rdtype == dns.rdatatype.PTR:
                    out.append((owner, rrset.ttl, 'PTR', _name(rd.target)))
                elif rrset.rdtype == dns.rdatatype.SOA:
                    assert (rd.serial, rd.refresh, rd.retry, rd.expire) == (0, 10, 10, 10)
                    host = _name(rd.mname)
                    assert _name(rd.rname) == ('hostmaster.' + host if host else 'hostmaster')
                    out.append((owner, rrset.ttl, 'SOA', host, rd.minimum))
                else:
                    raise AssertionError('unexpected rdtype %r' % rrset.rdtype)
        return tuple(out)
    info = {
        'id': m.id, 'qr': bool(m.flags & dns.flags.QR), 'aa': bool(m.flags & dns.flags.AA),
        'tc': bool(m.flags & dns.flags.TC), 'rd': bool(m.flags & dns.flags.RD),
        'ra': bool(m.flags & dns.flags.RA), 'edns': m.edns >= 0, 'payload': m.payload,
        'question': [(_name(q.name), q.rdtype) for q in m.question],
    }
    return m.rcode(), rrs(m.answer), rrs(m.authority), rrs(m.additional), info


def ref_semantic(options, pkt, seed=0, qidx=0):
    """oracle/binder_ref.py on a query packet -> (status, rcode, answers, authority, additional)."""
    labels, qtype, rd, opcode, _ = split_query(pkt)
    return R.on_query(options, labels, qtype, rd, seed, qidx, opcode).as_tuple()


def ref_options(snap_bytes, dns_domain, recursion=False, datacenter=''):
    zk = R.ZKCache(dns_domain).load_snapshot(snap_bytes.decode('utf-8').split('\n'))
    return R.Options(zk, dns_domain, datacenter, recursion)


# ---------------------------------------------------------------------------------------
# implementations under test, one calling convention
# ---------------------------------------------------------------------------------------
def make_impl(kind, dns_domain, snap_bytes, recursion=False, datacenter='', **kw):
    """kind 'oracle' -> oracle/liboracle.so;  kind 'gpu' -> the product through its C ABI."""
    if kind == 'oracle':
        from oracle_lib import Oracle
        return Oracle(dns_domain, datacenter, recursion, snapshot=snap_bytes)
    if kind == 'gpu':
        from binder_b200.engine import Engine
        return Engine(dns_domain, datacenter, recursion, snapshot=snap_bytes, **kw)
    raise ValueError(kind)


def resolve_list(impl, pkts, seed=0, qidx_base=0):
    """[packet bytes] -> [(status, response bytes)], plus the miss index list."""
    from binder_b200.synth import pack_batch
    data, off = pack_batch(pkts)
    out, out_off, out_len, status, miss = impl.resolve_batch(data, off, seed=seed, qidx_base=qidx_base)
    res = [(int(status[i]), bytes(out[out_off[i]:out_off[i] + out_len[i]])) for i in range(len(pkts))]
    return res, sorted(int(x) for x in miss)
