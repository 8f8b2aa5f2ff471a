// Reactive application state of the MegaScope front-end (counterpart of the refs in the reference's
// AppContent.vue): one plain object, topic-based subscriptions, helpers that fold the WebSocket
// "update" stream into per-(flag, layer) token-major tables.
export const FLAGS = {1:"QKV_mat_mul", 2:"RawAttentionScore_mat_mul", 3:"ContextLayer_mat_mul", 4:"MLP1_mat_mul",
                      5:"MLP2_mat_mul", 6:"Result", 7:"MLP2_Plot"};
export const FLAG_LABEL = {QKV_mat_mul:"QKV projections", RawAttentionScore_mat_mul:"attention probabilities",
  ContextLayer_mat_mul:"context layer", MLP1_mat_mul:"MLP fc1", MLP2_mat_mul:"MLP fc2", MLP2_Plot:"MLP2 PCA trajectory",
  Result:"output probabilities"};

const subs = new Map();
export const state = {
  mode:"inference", status:"disconnected", statusKind:"", running:false,
  meta:{num_layers:0, micro_batch_size:1, seq_length:0},
  promptTokens:[], genTokens:[], text:[], selToken:-1, layer:1, head:0, batch:0,
  updates:{},          // "flag:layer" -> [update messages in arrival order]
  probs:[],            // one entry per decode step: [[token, p], ...]
  history:[],          // finished runs kept for side-by-side comparison of perturbation experiments
};
export function on(topic, fn) { (subs.get(topic) || subs.set(topic, new Set()).get(topic)).add(fn); return () => subs.get(topic).delete(fn); }
export function emit(topic) { (subs.get(topic) || []).forEach(fn => fn(state)); (subs.get("*") || []).forEach(fn => fn(state, topic)); }
export function set(patch, topic="view") { Object.assign(state, patch); emit(topic); }

export function resetRun(meta) {
  state.meta = meta; state.updates = {}; state.probs = []; state.genTokens = []; state.text = [];
  state.promptTokens = (meta.prompt_tokens && meta.prompt_tokens[state.batch]) || (meta.prompt_tokens && meta.prompt_tokens[0]) || [];
  state.layer = Math.min(Math.max(1, state.layer), meta.num_layers || 1); state.running = true;
  emit("meta"); emit("data"); emit("tokens");
}
export function addUpdate(m) {
  if (m.update_type === 6) {
    if (Array.isArray(m.result)) state.probs.push(m.result);
    if (m.sampled !== undefined) state.genTokens.push(Array.isArray(m.sampled) ? m.sampled[state.batch] ?? m.sampled[0] : m.sampled);
    emit("probs"); emit("tokens"); return;
  }
  const k = m.update_type + ":" + m.layer_id;
  (state.updates[k] = state.updates[k] || []).push(m);
  emit("data");
}
export function finishRun(m) {
  state.running = false; if (m && m.text) state.text = m.text;
  state.history.unshift({when:new Date().toLocaleTimeString(), tokens:state.promptTokens.concat(state.genTokens), text:state.text});
  state.history.length = Math.min(state.history.length, 8);
  emit("tokens"); emit("history"); emit("data");
}
// flat helpers ------------------------------------------------------------------------------------
export const flat = x => Array.isArray(x) ? x.flat(Infinity) : [x];
export function shapeOf(e) { return e.args && e.args.length ? e.args : [flat(e.result).length]; }
// token-major rows of a vector-type capture (QKV / MLP1 / MLP2): every update carries [.., S, pixels];
// prefill + decode steps are concatenated along the token axis
export function tokenRows(flag, layer) {
  const rows = [];
  (state.updates[flag + ":" + layer] || []).forEach(e => {
    const shp = shapeOf(e), px = shp[shp.length - 1] || 1, f = flat(e.result);
    for (let i = 0; i + px <= f.length; i += px) rows.push(f.slice(i, i + px));
  });
  return rows;
}
export function latest(flag, layer) { const a = state.updates[flag + ":" + layer]; return a && a.length ? a[a.length - 1] : null; }
export function layersWith(flag) { return Object.keys(state.updates).filter(k => k.startsWith(flag + ":")).map(k => +k.split(":")[1]).sort((a, b) => a - b); }
