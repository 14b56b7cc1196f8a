// merkle's Hash and MerkleTree over libgenstark_b200.so: digests and trees stay in HBM (lib/Stark.ts:49-53,115,118,150,206;
// lib/components/LowDegreeProver.ts:45-46,52-53,86,109,116).
import { native, guarded, HASH_IDS } from './addon';
import { B200Matrix } from './B200Field';

export class B200Digests {
    constructor(readonly ctx: any, readonly handle: any) {}
    get length(): number { return native.digestsCount(this.handle); }
    toBuffers(): Buffer[] {
        const n = this.length, raw = Buffer.alloc(32 * n);
        guarded(() => native.digestsToBytes(this.ctx, this.handle, raw));
        return Array.from({ length: n }, (_, i) => raw.subarray(32 * i, 32 * i + 32));
    }
}

export class B200Hash {
    readonly digestSize = 32;
    readonly isOptimized = true;
    readonly id: number;
    constructor(readonly algorithm: 'sha256' | 'blake2s256', readonly ctx: any) {
        if (!(algorithm in HASH_IDS)) throw new TypeError(`Hash algorithm ${algorithm} is not supported`);
        this.id = HASH_IDS[algorithm];
    }
    digest(value: Buffer): Buffer { return guarded(() => native.hashDigest(this.id, value)); }
    merge(a: Buffer, b: Buffer): Buffer { return this.digest(Buffer.concat([a, b])); }
    digestValues(rows: B200Matrix, _valueSize?: number): B200Digests { return new B200Digests(this.ctx, guarded(() => native.hashDigestValues(this.ctx, this.id, rows.handle))); }
    mergeVectorRows(vectors: B200Matrix[]): B200Digests {
        return new B200Digests(this.ctx, guarded(() => native.hashMergeVectorRows(this.ctx, this.id, vectors.map(v => v.handle))));
    }
}

export interface BatchMerkleProof { values: Buffer[]; nodes: Buffer[][]; depth: number; }

export class B200MerkleTree {
    private constructor(readonly ctx: any, readonly handle: any, readonly hash: B200Hash) {}
    static create(values: B200Digests, hash: B200Hash): B200MerkleTree {
        return new B200MerkleTree(values.ctx, guarded(() => native.merkleCreate(values.ctx, hash.id, values.handle)), hash);
    }
    get root(): Buffer { return guarded(() => native.merkleRoot(this.ctx, this.handle)); }
    /** wire layout of the blob: u8 depth, u16 count, count x 32-byte leaf, then per leaf: u8 n, n x 32-byte node */
    proveBatch(indexes: number[]): BatchMerkleProof {
        const out = Buffer.alloc(1 << 20);
        const len: number = guarded(() => native.merkleProveBatch(this.ctx, this.handle, Uint32Array.from(indexes), indexes.length, out));
        return parseBatchProof(out.subarray(0, len));
    }
    static verifyBatch(root: Buffer, indexes: number[], proof: BatchMerkleProof, hash: B200Hash): boolean {
        return guarded(() => native.merkleVerifyBatch(hash.id, root, Uint32Array.from(indexes), indexes.length, writeBatchProof(proof))) === 1;
    }
}

export function parseBatchProof(b: Buffer): BatchMerkleProof {
    let o = 0;
    const depth = b.readUInt8(o); o += 1;
    const count = b.readUInt16LE(o); o += 2;
    const values: Buffer[] = [], nodes: Buffer[][] = [];
    for (let i = 0; i < count; i++, o += 32) values.push(b.subarray(o, o + 32));
    for (let i = 0; i < count; i++) {
        const n = b.readUInt8(o); o += 1;
        const col: Buffer[] = [];
        for (let j = 0; j < n; j++, o += 32) col.push(b.subarray(o, o + 32));
        nodes.push(col);
    }
    return { values, nodes, depth };
}
export function writeBatchProof(p: BatchMerkleProof): Buffer {
    const parts: Buffer[] = [Buffer.from([p.depth]), Buffer.from([p.values.length & 255, p.values.length >> 8]), ...p.values];
    for (const col of p.nodes) parts.push(Buffer.from([col.length]), ...col);
    return Buffer.concat(parts);
}
