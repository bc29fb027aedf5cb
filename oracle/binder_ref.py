"""TEST INFRASTRUCTURE — literal, semantic-level CPU restatement of binder's resolve path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this
module.  It is NOT the product and the product never routes through it.

What it restates (file:line relative to /root/reference):
  * lib/zk.js:20-48,55-67   ZKCache   (lookup / reverseLookup / isReady)
  * lib/zk.js:78-119        TreeNode  (name keeps case, domain lower-cased, children = insertion order)
  * lib/zk.js:139-194       TreeNode.onDataChanged (JSON ingest, reverse-map maintenance)
  * lib/zk.js:225-228       domainToPath
  * lib/server.js:40-53     shuffle (Fisher-Yates; Math.random() replaced by the seeded
                            counter RNG of DESIGN.md "Shuffle RNG")
  * lib/server.js:55-65     isSuffix / stripSuffix
  * lib/server.js:67-134    resolvePtr
  * lib/server.js:136-429   resolve
  * lib/server.js:491-506   onQuery type dispatch

It works on *decoded* queries (name string, qtype, RD flag) and returns a semantic
Response (status, rcode, answer / authority / additional RR tuples).  The byte-level
codec (mname@1.5.1, not vendored in the reference) is restated in oracle/oracle.cpp;
tests decode that oracle's bytes with dnspython and compare against this module, so
the codec is checked by an independent DNS implementation and the resolve logic by two
independent restatements (this file: Python dicts; oracle.cpp: C++).

Pure-Python loops: small cases only.
"""
import json
import re

# status codes shared with include/binder_b200.h
ANSWERED, MISS_RECURSE, DROPPED = 0, 1, 2

RCODE = {'noerror': 0, 'eserver': 2, 'servfail': 2, 'nxdomain': 3, 'enotimp': 4,
         'refused': 5}
NOTIMP_DEFAULT = 4          # mname's rcode when nothing was set and nothing was added
                            # (lib/server.js:346-350 comment)

QTYPE_NAMES = {1: 'A', 12: 'PTR', 33: 'SRV'}

HOSTLIKE = ('db_host', 'host', 'load_balancer', 'moray_host', 'redis_host',
            'ops_host', 'rr_host')                 # lib/server.js:302-308, lib/zk.js:173-179
SERVICE_KID_TYPES = ('load_balancer', 'moray_host', 'ops_host', 'rr_host',
                     'redis_host')                 # lib/server.js:355-359

_UNDEF = object()           # JS `undefined`


def _get(obj, key):
    """JS property read on a JSON.parse()d value (own properties only)."""
    if isinstance(obj, dict):
        return obj.get(key, _UNDEF)
    return _UNDEF


def _is_object(v):
    """typeof v === 'object' && v !== null."""
    return isinstance(v, (dict, list))


def fmix32(x):
    x &= 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x85EBCA6B) & 0xFFFFFFFF
    x ^= x >> 13
    x = (x * 0xC2B2AE35) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def shuffle_rand(seed, qidx, i):
    """Counter-based replacement for Math.random() at lib/server.js:46.

    Returns j in [0, i] = floor(u * (i + 1)) with u = r / 2**32."""
    lo = seed & 0xFFFFFFFF
    hi = (seed >> 32) & 0xFFFFFFFF
    r = fmix32(fmix32(fmix32(lo) ^ hi ^ ((qidx * 0x9E3779B1) & 0xFFFFFFFF))
               + ((i * 0x85EBCA77) & 0xFFFFFFFF))
    return (r * (i + 1)) >> 32


def shuffle(arr, seed, qidx):
    """lib/server.js:40-53."""
    if len(arr) == 0:
        return arr
    i = len(arr)
    while True:
        i -= 1
        if not i > 0:
            break
        j = shuffle_rand(seed, qidx, i)
        arr[i], arr[j] = arr[j], arr[i]
    return arr


def is_suffix(suffix, s):
    """lib/server.js:55-58."""
    idx = s.rfind(suffix)
    return idx >= 0 and idx + len(suffix) == len(s)


def valid_ipv4(s):
    """Contract: an address the A-record encoder accepts — dotted quad, decimal,
    each octet "0" or [1-9][0-9]{0,2} with value <= 255 (no leading zeros)."""
    if not isinstance(s, str):
        return False
    parts = s.split('.')
    if len(parts) != 4:
        return False
    for p in parts:
        if not (1 <= len(p) <= 3) or not all('0' <= c <= '9' for c in p):
            return False
        if (len(p) > 1 and p[0] == '0') or int(p) > 255:
            return False
    return True


def valid_uint(v, limit):
    """Contract for ttl / port numbers: JSON number that is a non-negative integer
    below `limit` (bool is not a number in JSON)."""
    if isinstance(v, bool):
        return False
    if isinstance(v, int):
        return 0 <= v < limit
    if isinstance(v, float):
        return v == int(v) and 0 <= v < limit
    return False


TTL_LIMIT = 1 << 31
PORT_LIMIT = 1 << 16


def url_hostname(primary):
    """url.parse(primary).hostname for the 'scheme://[user[:pw]@]host[:port]/path'
    shapes binder's database records use (lib/server.js:297-298)."""
    if not isinstance(primary, str):
        return None
    m = re.match(r'^[A-Za-z][A-Za-z0-9+.-]*://([^/?#]*)', primary)
    if not m:
        return None
    auth = m.group(1)
    if '@' in auth:
        auth = auth.rsplit('@', 1)[1]
    host = auth.rsplit(':', 1)[0] if ':' in auth else auth
    return host.lower()


class TreeNode(object):
    """lib/zk.js:78-119."""

    def __init__(self, cache, p_domain, name):
        # strings are handled as the latin-1 view of their UTF-8 bytes, so that they compare
        # byte-for-byte with the latin-1 view of wire names
        name = name.encode('utf-8', 'surrogatepass').decode('latin-1')
        self.tn_name = name
        dom = name
        if len(p_domain) > 0:
            dom += '.' + p_domain
        self.tn_domain = ascii_lower(dom)
        self.tn_kids = {}            # insertion ordered (py3.7+), like Object.keys
        self.tn_data = None
        self.tn_ip = None
        self.tn_cache = cache
        cache.ca_treeNodes[self.tn_domain] = self

    @property
    def children(self):
        return list(self.tn_kids.values())

    def on_data_changed(self, raw=None, parsed=_UNDEF):
        """lib/zk.js:139-194.  `raw` = znode bytes as str; or pass the already
        JSON.parse()d value in `parsed`."""
        if parsed is _UNDEF:
            try:
                parsed = json.loads(raw, parse_constant=_no_constants)
            except ValueError:
                return                              # :144-148 + :149 (undefined)
        if not (parsed is None or isinstance(parsed, (dict, list))):
            return                                  # :149-154 typeof !== 'object'
        self.tn_data = parsed
        t = _get(parsed, 'type')
        if parsed is None or not isinstance(t, str):
            return                                  # :157-165
        if t in HOSTLIKE:
            record = _get(parsed, t)
            if not _is_object(record):
                return                              # :181-182
            addr = _get(record, 'address')
            if self.tn_ip:
                self.tn_cache.ca_revLookup.pop(self.tn_ip, None)
            self.tn_ip = None
            # contract: only string addresses index the reverse map
            if isinstance(addr, str) and addr:
                key = addr.encode('utf-8', 'surrogatepass').decode('latin-1')
                self.tn_ip = key
                self.tn_cache.ca_revLookup[key] = self


def _no_constants(name):
    raise ValueError('JSON.parse rejects ' + name)


def ascii_lower(s):
    return ''.join(chr(ord(c) + 32) if 'A' <= c <= 'Z' else c for c in s)


def domain_to_path(domain):
    """lib/zk.js:225-228."""
    return '/' + '/'.join(reversed(domain.split('.')))


class ZKCache(object):
    """lib/zk.js:20-76 read side, filled from a snapshot instead of ZK watchers."""

    def __init__(self, domain):
        self.ca_treeNodes = {}
        self.ca_revLookup = {}
        self.ca_domain = domain
        self.by_path = {}

    def unbind(self, node):
        """lib/zk.js:195-208: the subtree leaves ca_treeNodes (a key only if it is still the node's
        own); ca_revLookup is never touched."""
        for kid in node.tn_kids.values():
            self.unbind(kid)
        if self.ca_treeNodes.get(node.tn_domain) is node:
            del self.ca_treeNodes[node.tn_domain]

    def _forget_paths(self, node, path):
        for name, kid in node.tn_kids.items():
            self._forget_paths(kid, path + '/' + name)
        self.by_path.pop(path, None)

    def apply_delta(self, lines):
        """Watch events on a loaded cache, as JSON lines: a known path = dataChanged (:139-194), a new
        path = a child appended by childrenChanged (:120-130) plus its data, {"path": P, "deleted":
        true} = the child vanishing from its parent's list (:131-133 -> unbind)."""
        return self._apply(lines, True)

    def load_snapshot(self, lines):
        """Snapshot = JSON lines {"path": "/com/foo/x", "data": <value>} (or "raw":
        "<znode bytes>"), parents before children, children in ZK child-list order.
        Mirrors rebuildCache (:68-76) + the watcher callbacks."""
        parts = self.ca_domain.split('.')
        root = TreeNode(self, '.'.join(parts[1:]), parts[0])
        self.by_path[domain_to_path(self.ca_domain)] = root
        return self._apply(lines, False)

    def _apply(self, lines, allow_delete):
        root_path = domain_to_path(self.ca_domain)
        root = self.by_path[root_path]
        for line in lines:
            if isinstance(line, (bytes, bytearray)):
                line = line.decode('utf-8')
            line = line.strip()
            if not line:
                continue
            ent = json.loads(line)
            path = ent['path']
            deleting = ent.get('deleted') is True
            if deleting and not allow_delete:
                raise ValueError('a snapshot states what exists')
            if path == root_path:
                if deleting:
                    raise ValueError('the root of the mirrored subtree stays')
                node = root
            else:
                ppath, _, name = path.rpartition('/')
                parent = self.by_path.get(ppath)
                if parent is None or name == '':
                    continue                        # outside the watched subtree
                node = parent.tn_kids.get(name)
                if deleting:
                    if node is not None:
                        self.unbind(node)
                        self._forget_paths(node, path)
                        del parent.tn_kids[name]
                    continue
                if node is None:
                    node = TreeNode(self, parent.tn_domain, name)
                    parent.tn_kids[name] = node
                    self.by_path[path] = node
            if 'raw' in ent:
                node.on_data_changed(raw=ent['raw'])
            elif 'data' in ent:
                node.on_data_changed(parsed=ent['data'])
        return self

    def isReady(self):
        return self.ca_domain in self.ca_treeNodes

    def lookup(self, domain):
        return self.ca_treeNodes.get(domain)

    def reverseLookup(self, ip):
        return self.ca_revLookup.get(ip)


class Response(object):
    def __init__(self):
        self.status = ANSWERED
        self.rcode = None            # None = nothing set yet
        self.answers = []            # (owner, ttl, 'A', addr) / (owner, ttl, 'SRV', port, target)
        self.authority = []          # (owner, ttl, 'SOA', host, minimum)
        self.additional = []         # (owner, ttl, 'A', addr)

    def setError(self, name):
        self.rcode = RCODE[name]

    def addAnswer(self, rr):
        self.answers.append(rr)

    def final_rcode(self):
        if self.rcode is not None:
            return self.rcode
        return 0 if self.answers else NOTIMP_DEFAULT

    def as_tuple(self):
        return (self.status, self.final_rcode(), tuple(self.answers), tuple(self.authority),
                tuple(self.additional))


# JS /^(_[^_.]*)[.](_[^_.]*)[.](.*)/ : '.' does not match \n or \r (nor U+2028/9, which
# cannot occur in the latin-1 view of a wire name), and there is no '$' anchor, so group 3
# silently stops at the first line terminator.
_SRV_RE = re.compile(r'^(_[^_.]*)[.](_[^_.]*)[.]([^\n\r]*)')


class Options(object):
    def __init__(self, zkCache, dnsDomain, datacenterName='', recursion=False):
        self.zkCache = zkCache
        self.dnsDomain = dnsDomain
        self.datacenterName = datacenterName
        self.recursion = recursion
        self.recursion_filter = None        # (region dnsDomain, forwardable dc names, ptr forwardable) or None


def recursion_forwards(flt, domain, is_ptr):
    """lib/recursion.js:329-344,377-379: would Recursion.resolve() forward this miss anywhere?  `domain` is
    query.name() as received."""
    dns_domain, dcs, ptr = flt
    if is_ptr:
        return bool(ptr)                                        # :346-354: every datacenter's resolvers
    if domain.find(dns_domain, max(len(domain) - len(dns_domain), 0)) == -1:
        return False                                            # :330-333
    p = domain[:max(len(domain) - len(dns_domain) - 1, 0)]     # :338-339
    dc = p[p.rfind('.') + 1:]                                   # :340
    return dc in dcs                                            # :341-343


def encodable(name):
    """Every '.'-separated label fits the wire format (1..63 bytes, <= 255 total)."""
    if name == '':
        return True
    labs = name.encode('latin-1').split(b'.')
    if any(len(l) < 1 or len(l) > 63 for l in labs):
        return False
    return sum(len(l) + 1 for l in labs) + 1 <= 255


def _record_ttl(record, default=30):
    """lib/server.js:270-274 / 124-128.  Returns (ttl, ok)."""
    ttl = default
    ok = True
    v = _get(record, 'ttl')
    if v is not _UNDEF:
        ttl = v
    sub = _get(record, _get(record, 'type'))
    v = _get(sub, 'ttl')
    if v is not _UNDEF:
        ttl = v
    if not valid_uint(ttl, TTL_LIMIT):
        ok = False
    return (int(ttl) if ok else 0), ok


def resolve_ptr(options, name, rd, resp):
    """lib/server.js:67-134."""
    parts = list(reversed(name.split('.')))
    if len(parts) < 2 or parts[0] != 'arpa' or parts[1] != 'in-addr':
        resp.setError('refused')
        return resp
    ip = '.'.join(parts[2:])
    zk = options.zkCache
    if zk is None or not zk.isReady():
        resp.setError('eserver')
        return resp
    node = zk.reverseLookup(ip)
    if not node:
        if options.recursion and rd:
            if options.recursion_filter and not recursion_forwards(options.recursion_filter, name, True):
                resp.setError('refused')
                return resp
            resp.status = MISS_RECURSE
            return resp
        resp.setError('refused')
        return resp
    record = node.tn_data
    ttl, ok = _record_ttl(record)
    if not ok or not encodable(node.tn_domain):
        resp.setError('servfail')       # contract deviation: reference would throw in mname
        return resp
    resp.addAnswer((name, ttl, 'PTR', node.tn_domain))
    return resp


def resolve(options, name, qtype, rd, resp, seed, qidx):
    """lib/server.js:136-429."""
    domain = name
    service = protocol = None
    srvmatch = _SRV_RE.match(domain)
    if qtype == 'SRV':
        if not srvmatch or len(srvmatch.group(3)) < 1:
            resp.setError('refused')
            return resp
        service, protocol, domain = srvmatch.group(1), srvmatch.group(2), srvmatch.group(3)

    if options.dnsDomain:
        if not is_suffix('.' + options.dnsDomain, domain):
            resp.setError('refused')
            return resp
        # :167-175 "doubled-up suffix" — dead code: stripSuffix() appends '...', so
        # neither isSuffix() can ever be true.  Deliberately not restated as a refusal.

    zk = options.zkCache
    if zk is None or not zk.isReady():
        resp.setError('eserver')
        return resp

    if len(domain) < 1:
        resp.setError('refused')
        return resp

    domain = ascii_lower(domain)
    if re.search(r'[^a-z0-9_.-]', domain):
        resp.setError('refused')
        return resp

    node = zk.lookup(domain)
    if not node:
        if options.recursion and rd:
            if options.recursion_filter and not recursion_forwards(options.recursion_filter, name, False):
                resp.setError('refused')
                return resp
            resp.status = MISS_RECURSE
            return resp
        resp.setError('refused')
        return resp

    record = node.tn_data
    rtype = _get(record, 'type')
    if (not record) or not isinstance(rtype, str) or not _is_object(_get(record, rtype)):
        resp.setError('servfail')
        return resp

    ttl, ok = _record_ttl(record)
    if not ok:
        resp.setError('servfail')       # contract deviation (non-integer ttl)
        return resp

    if service is not None and rtype != 'service':
        resp.setError('noerror')
        resp.authority.append((domain, ttl, 'SOA', options.dnsDomain, ttl))
        return resp

    if rtype == 'database':
        addr = url_hostname(_get(_get(record, 'database'), 'primary'))
        if not valid_ipv4(addr):
            resp.setError('servfail')   # contract deviation (reference: ARecord throws)
            return resp
        resp.addAnswer((domain, ttl, 'A', addr))
    elif rtype in HOSTLIKE:
        addr = _get(_get(record, rtype), 'address')
        if not valid_ipv4(addr):
            resp.setError('servfail')   # contract deviation
            return resp
        resp.addAnswer((domain, ttl, 'A', addr))
    elif rtype == 'service':
        s = _get(record, 'service')
        inner = _get(s, 'service')
        if inner is None:
            resp.setError('servfail')   # contract deviation (reference: TypeError on null.ttl)
            return resp
        if _is_object(inner):
            s = inner
        v = _get(s, 'ttl')
        if v is not _UNDEF:
            if not valid_uint(v, TTL_LIMIT):
                resp.setError('servfail')
                return resp
            ttl = int(v)
        if service is not None and (service != _get(s, 'srvce') or protocol != _get(s, 'proto')):
            resp.setError('nxdomain')
            return resp
        resp.setError('noerror')
        kids = [k for k in node.children
                if k.tn_data and _get(k.tn_data, 'type') in SERVICE_KID_TYPES]
        kids = shuffle(kids, seed, qidx)
        for knode in kids:
            krec = knode.tn_data
            ktype = _get(krec, 'type')
            ksub = _get(krec, ktype)
            if not _is_object(ksub):
                resp.setError('eserver')
                break
            a = _get(ksub, 'address')
            if a is None:
                continue
            ports = _get(ksub, 'ports')
            if ports is _UNDEF or (isinstance(ports, list) and len(ports) < 1):
                ports = [_get(s, 'port')]
            rttl = ttl
            v = _get(krec, 'ttl')
            if v is not _UNDEF:
                rttl = v
            v = _get(ksub, 'ttl')
            if v is not _UNDEF:
                rttl = v
            # contract deviations: anything mname's record constructors would throw on
            # is treated like the reference's own "bad zk info" branch (:366-376)
            bad = (not valid_ipv4(a)) or (not valid_uint(rttl, TTL_LIMIT))
            if service is not None:
                bad = bad or not isinstance(ports, list) or \
                    any(not valid_uint(p, PORT_LIMIT) for p in ports) or \
                    not encodable(knode.tn_name + '.' + domain)
            if bad:
                resp.setError('eserver')
                break
            rttl = int(rttl)
            if service is not None:
                nm = knode.tn_name + '.' + domain
                for p in ports:
                    resp.addAnswer((name, ttl, 'SRV', int(p), nm))
                resp.additional.append((nm, rttl, 'A', a))
            else:
                if ttl < rttl:
                    rttl = ttl
                resp.addAnswer((domain, rttl, 'A', a))
    else:
        pass                            # :419-424 unknown type: nothing added, nothing set
    return resp


def on_query(options, labels, qtype_num, rd, seed=0, qidx=0, opcode=0):
    """lib/server.js:471-507 dispatch.  `labels` = the QNAME's wire labels (bytes each);
    query.name() is their latin-1 view joined by '.', '' for the root."""
    resp = Response()
    t = QTYPE_NAMES.get(qtype_num)
    if t is None or opcode != 0:
        resp.setError('enotimp')
        return resp
    # contract deviation (DESIGN.md "in-label dots"): a label containing a literal '.'
    # cannot be told apart from two labels once mname joins the name with '.', so such a
    # QNAME is refused before any lookup.
    if any(b'.' in l for l in labels):
        resp.setError('refused')
        return resp
    name = '.'.join(l.decode('latin-1') for l in labels)
    if t in ('A', 'SRV'):
        return resolve(options, name, t, rd, resp, seed, qidx)
    return resolve_ptr(options, name, rd, resp)
