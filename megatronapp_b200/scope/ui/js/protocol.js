// WebSocket client of the MegaScope servers (inference: megatronapp_b200.inference.server.InferenceWSServer,
// training: megatronapp_b200.scope.training_server) -- same JSON protocol as the reference
// (tools/run_text_generation_server / ws server :5000): ping/pong heart-beat, "generate" / "run_training_step"
// requests carrying visualization_flags + disturbance_configs + compressor_config, and the start / update /
// finish / error stream coming back.
import {state, set, resetRun, addUpdate, finishRun} from "./store.js";

let ws = null, beat = null, retries = 0;
export function connect(url) {
  if (ws) { ws.onclose = null; ws.close(); }
  set({status:"connecting...", statusKind:""}, "status");
  ws = new WebSocket(url);
  ws.onopen = () => { retries = 0; set({status:"connected", statusKind:"ok"}, "status"); send({type:"ping"});
    clearInterval(beat); beat = setInterval(() => send({type:"ping"}), 25000); };
  ws.onclose = ev => { clearInterval(beat); set({status: ev.wasClean ? "disconnected" : "connection lost", statusKind: ev.wasClean ? "" : "err", running:false}, "status");
    if (!ev.wasClean && retries++ < 3) setTimeout(() => connect(url), 3000); };
  ws.onerror = () => set({status:"connection error", statusKind:"err"}, "status");
  ws.onmessage = ev => { let m; try { m = JSON.parse(ev.data); } catch (e) { set({status:"unparsable message", statusKind:"err"}, "status"); return; }
    if (m.type === "start") { resetRun(m); set({status:"running", statusKind:"ok"}, "status"); }
    else if (m.type === "update") addUpdate(m);
    else if (m.type === "finish" || m.type === "finish-update") { finishRun(m); set({status:"finished", statusKind:"ok"}, "status"); }
    else if (m.type === "error") set({status:m.message || "server error", statusKind:"err", running:false}, "status");
    else if (m.type !== "pong") console.warn("unknown message", m); };
}
export function connected() { return ws && ws.readyState === 1; }
export function send(obj) { if (connected()) ws.send(JSON.stringify(obj)); }
export function run(request) {
  if (!connected()) { set({status:"not connected", statusKind:"err"}, "status"); return; }
  send(state.mode === "training" ? {type:"run_training_step", ...request.configs} : {type:"generate", ...request.sampling, ...request.configs});
  set({running:true, status:"request sent", statusKind:""}, "status");
}
