// Stark.prove() in one crossing: gs_stark_prove returns the serialized proof (lib/Serializer.ts:35-79), which the reference's own
// Serializer.parseProof turns back into a StarkProof -- so lib/Stark.ts keeps its API (prove returns StarkProof) and verify() is untouched.
import { native, guarded, HASH_IDS, packElements } from './addon';

export interface Assertion { step: number; register: number; value: bigint; }

export class B200Prover {
    readonly ctx: any;
    readonly stark: any;
    /** airBlob: the flattened AirAssembly module (transition / evaluation procedures, static registers); layout = genstark_b200/air.py pack_air */
    constructor(airBlob: Buffer, hashAlgorithm: string, exeQueryCount: number, friQueryCount: number, device = 0) {
        this.ctx = guarded(() => native.ctxCreate(device));
        this.stark = guarded(() => native.starkCreate(this.ctx, airBlob, HASH_IDS[hashAlgorithm], exeQueryCount, friQueryCount));
    }
    /** assertions: lib/Stark.ts:81-87; initState: first trace row; inputTraces: n_input x T elements (16 bytes LE each) or null; shapes: iShapes */
    prove(assertions: Assertion[], initState: bigint[], inputTraces: Buffer | null, shapes: number[][]): Buffer {
        if (!Array.isArray(assertions)) throw new TypeError('Assertions parameter must be an array');
        if (assertions.length === 0) throw new TypeError('At least one assertion must be provided');
        const a = Buffer.alloc(24 * assertions.length);
        assertions.forEach((x, i) => { a.writeUInt32LE(x.register, 24 * i); a.writeUInt32LE(x.step, 24 * i + 4); packElements([x.value]).copy(a, 24 * i + 8); });
        return guarded(() => native.starkProve(this.stark, a, assertions.length, packElements(initState), inputTraces, packShapes(shapes)));
    }
    /** the stage log the reference prints through its Logger (lib/Stark.ts:92-160), as data */
    stageTimes(): [string, number][] { return JSON.parse(native.starkStageTimes(this.stark) || '[]'); }
    destroy(): void { native.starkDestroy(this.stark); native.ctxDestroy(this.ctx); }
}

export function packShapes(shapes: number[][]): Buffer {
    const parts: Buffer[] = [Buffer.from([shapes.length])];
    for (const s of shapes) {
        const b = Buffer.alloc(1 + 4 * s.length);
        b.writeUInt8(s.length, 0);
        s.forEach((x, i) => b.writeUInt32LE(x, 1 + 4 * i));
        parts.push(b);
    }
    return Buffer.concat(parts);
}
