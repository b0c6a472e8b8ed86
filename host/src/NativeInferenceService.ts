// Drop-in replacement for client/src/services/OllamaService.ts: same 8-method surface, same
// InferenceRequest / InferenceResponse / StreamResponse types (client/src/types/index.ts), but every call goes
// to the in-process native engine (host/napi/addon.cc -> libgridllm_native.so) instead of HTTP to Ollama.
// WorkerClientService.ts needs one line changed: `new OllamaService()` -> `new NativeInferenceService()`
// (WorkerClientService.ts:32).  Not executable in the build image (no node); the Python twin
// gridllm_b200/service.py is what the tests run.
import { config } from "@/config";
import { logger } from "@/utils/logger";
import { OllamaModel, InferenceRequest, InferenceResponse, StreamResponse } from "@/types";
import * as fs from "fs";

// eslint-disable-next-line @typescript-eslint/no-var-requires
const native = require("../napi/build/Release/gridllm_native.node");

interface NativeStats {
	promptEvalCount: number; evalCount: number; promptEvalDurationNs: number; evalDurationNs: number;
	totalDurationNs: number; loadDurationNs: number; doneReason: number;
}

export class NativeInferenceService {
	private engines = new Map<string, unknown>();
	private isConnected = false;
	private lastHealthCheck = new Date();
	// model name -> GGUF path, e.g. GRIDLLM_MODELS="llama3:8b=/models/llama3-8b-q4_K_M.gguf"
	private models: Record<string, string> = Object.fromEntries(
		(process.env.GRIDLLM_MODELS || "").split(",").filter(Boolean).map((kv) => kv.split("=") as [string, string])
	);
	private device = parseInt(process.env.GRIDLLM_DEVICE || "0", 10);

	private engine(name: string): unknown {
		if (!this.models[name]) throw new Error(`model '${name}' not found`);
		if (!this.engines.has(name)) this.engines.set(name, native.createEngine(this.models[name], this.device, {}));
		return this.engines.get(name);
	}

	async checkHealth(): Promise<boolean> {                       // OllamaService.ts:65-83
		this.isConnected = native.deviceCount() > this.device;
		this.lastHealthCheck = new Date();
		return this.isConnected;
	}

	async getAvailableModels(): Promise<OllamaModel[]> {           // :85-95
		return Object.entries(this.models).map(([name, path]) => {
			const st = fs.statSync(path);
			return { name, digest: `${st.size}-${st.mtimeMs}`, size: st.size, modified_at: st.mtime.toISOString(),
				details: { format: "gguf", family: "llama", families: ["llama"], parameter_size: "", quantization_level: "" } };
		});
	}

	async validateModel(modelName: string): Promise<boolean> {     // :340-351, O(1) instead of GET /api/tags per job
		return !!this.models[modelName];
	}

	private ids(e: unknown, request: InferenceRequest): Int32Array {
		const pre = request.metadata?.prompt_token_ids;
		if (pre) return Int32Array.from(pre);
		const ctx: number[] | undefined = request.metadata?.context;          // conversation so far (OllamaService.ts:224-226)
		if (ctx?.length) return Int32Array.from([...ctx, ...native.tokenize(e, request.prompt || "", false, false)]);
		return native.tokenize(e, request.prompt || "", true, false);
	}

	// options.stop: same rule as gridllm_b200/service.py::StopFilter -- end at the first stop string in the generated text,
	// hold back text that could still become one.  feed() returns the text a token releases; hit ends the native call.
	private stopFilter(stops: string[]) {
		const st = { text: "", held: "", hit: false,
			feed(piece: string): string {
				if (st.hit) return "";
				const buf = st.held + piece;
				const cuts = stops.map((s) => buf.indexOf(s)).filter((i) => i >= 0);
				if (cuts.length) { st.hit = true; st.held = ""; const out = buf.slice(0, Math.min(...cuts)); st.text += out; return out; }
				let keep = 0;
				for (const s of stops) for (let n = Math.min(s.length - 1, buf.length); n > 0; --n) if (buf.endsWith(s.slice(0, n))) { keep = Math.max(keep, n); break; }
				const out = buf.slice(0, buf.length - keep); st.held = buf.slice(buf.length - keep); st.text += out; return out;
			},
			flush(): string { const out = st.hit ? "" : st.held; st.held = ""; st.text += out; return out; } };
		return st;
	}

	private toResponse(request: InferenceRequest, text: string, ids: Int32Array, st: NativeStats): InferenceResponse {
		return { id: request.id, model: request.model, created_at: new Date().toISOString(), response: text, done: true,
			done_reason: st.doneReason === 0 ? "stop" : "length", total_duration: st.totalDurationNs, load_duration: st.loadDurationNs,
			prompt_eval_count: st.promptEvalCount, prompt_eval_duration: st.promptEvalDurationNs, eval_count: st.evalCount,
			eval_duration: st.evalDurationNs, context: Array.from(ids), system_fingerprint: "fp_gridllm_b200_native" };
	}

	// InferenceRequest.options -> gl_sample_opts; temperature absent / 0 = greedy, a sampled request without seed draws one
	private sampleOpts(request: InferenceRequest) {
		const o = request.options ?? {};
		const temperature = o.temperature ?? 0;
		return { numPredict: o.num_predict || 128, ignoreEos: !!o.ignore_eos, temperature, topK: o.top_k ?? 0, topP: o.top_p ?? 1,
			seed: BigInt(o.seed ?? (temperature > 0 ? Math.floor(Math.random() * 2 ** 53) : 0)) };
	}

	async generateResponse(request: InferenceRequest): Promise<InferenceResponse> {   // :97-184
		try {
			const e = this.engine(request.model);
			const stop = request.options?.stop;
			const stops: string[] = (typeof stop === "string" ? [stop] : stop || []).filter(Boolean);
			if (!stops.length) {
				const out = await native.generate(e, this.ids(e, request), this.sampleOpts(request), null);
				return this.toResponse(request, native.detokenize(e, out.ids), out.ids, out.stats);
			}
			const f = this.stopFilter(stops);         // a non-zero return of the token callback cancels gl_generate
			const out = await native.generate(e, this.ids(e, request), this.sampleOpts(request),
				(_id: number, _lp: number, piece: string) => { f.feed(piece); return f.hit; });
			f.flush();
			const res = this.toResponse(request, f.text, out.ids, out.stats);
			if (f.hit) res.done_reason = "stop";
			return res;
		} catch (error) {
			throw new Error(`Inference failed: ${error instanceof Error ? error.message : "Unknown error"}`);
		}
	}

	async *generateStreamResponse(request: InferenceRequest): AsyncGenerator<StreamResponse> {   // :186-284
		try {
			const e = this.engine(request.model);
			const queue: StreamResponse[] = [];
			let wake: (() => void) | null = null;
			const done = native.generate(e, this.ids(e, request), this.sampleOpts(request),
				(_id: number, _lp: number, piece: string) => { queue.push({ id: request.id, response: piece, done: false }); wake?.(); });
			let finished = false;
			done.then(() => { finished = true; wake?.(); }, () => { finished = true; wake?.(); });
			while (!finished || queue.length) {
				if (queue.length) { yield queue.shift()!; continue; }
				await new Promise<void>((r) => (wake = r));
			}
			await done;
			yield { id: request.id, response: "", done: true };
		} catch (error) {
			throw new Error(`Streaming inference failed: ${error instanceof Error ? error.message : "Unknown error"}`);
		}
	}

	async generateChatResponse(request: InferenceRequest): Promise<InferenceResponse> {   // :353-449
		if (!request.metadata?.messages) throw new Error("Chat inference failed: Chat request must include messages in metadata");
		const prompt = request.metadata.messages.map((m) => `<|start_header_id|>${m.role}<|end_header_id|>\n\n${m.content}<|eot_id|>`).join("")
			+ "<|start_header_id|>assistant<|end_header_id|>\n\n";
		const r = await this.generateResponse({ ...request, prompt });
		const { response, ...rest } = r;
		return { ...rest, message: { role: "assistant", content: response } };
	}

	async *generateChatStreamResponse(request: InferenceRequest): AsyncGenerator<StreamResponse> {   // :451-599
		if (!request.metadata?.messages) throw new Error("Chat streaming inference failed: Chat request must include messages in metadata");
		const prompt = request.metadata.messages.map((m) => `<|start_header_id|>${m.role}<|end_header_id|>\n\n${m.content}<|eot_id|>`).join("")
			+ "<|start_header_id|>assistant<|end_header_id|>\n\n";
		yield* this.generateStreamResponse({ ...request, prompt });
	}

	async generateEmbedding(request: InferenceRequest): Promise<InferenceResponse> {   // :601-665
		try {
			if (!request.input) throw new Error("Input is required for embedding requests");
			const e = this.engine(request.model);
			const texts = Array.isArray(request.input) ? request.input : [request.input];
			const seqs = texts.map((t) => native.tokenize(e, t, true, false) as Int32Array);
			const offsets = new Int32Array(seqs.length + 1);
			seqs.forEach((s, i) => (offsets[i + 1] = offsets[i] + s.length));
			const flat = new Int32Array(offsets[seqs.length]);
			seqs.forEach((s, i) => flat.set(s, offsets[i]));
			const out = await native.embed(e, flat, offsets);
			const dim = out.embeddings.length / seqs.length;
			return { id: request.id, model: request.model,
				embeddings: seqs.map((_, i) => Array.from(out.embeddings.subarray(i * dim, (i + 1) * dim))),
				total_duration: out.stats.totalDurationNs, load_duration: out.stats.loadDurationNs, prompt_eval_count: out.stats.promptEvalCount };
		} catch (error) {
			throw new Error(`Embedding failed: ${error instanceof Error ? error.message : "Unknown error"}`);
		}
	}

	getConnectionStatus() { return { isConnected: this.isConnected, lastHealthCheck: this.lastHealthCheck }; }
	async pullModel(modelName: string): Promise<void> { throw new Error(`Failed to pull model ${modelName}: local GGUF files only`); }
	async deleteModel(modelName: string): Promise<void> { throw new Error(`Failed to delete model ${modelName}: not managed`); }
}

export default NativeInferenceService;
void config; void logger;
