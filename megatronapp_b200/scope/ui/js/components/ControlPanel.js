// Left column: connection, prompts + sampling, capture switches, compressor, perturbation config, run button
// (reference AppContent.vue control collapse items + "Perturbation Config").
import {Component, h, panel} from "./base.js";
import {state, set, FLAGS, FLAG_LABEL} from "../store.js";
import {connect, run} from "../protocol.js";

const inp = (type, value, extra = {}) => h("input", {type, value, ...extra});
export class ControlPanel extends Component {
  constructor() {
    super("section");
    this.url = inp("text", "ws://127.0.0.1:5000"); this.mode = h("select", {}, h("option", {value:"inference"}, "inference"), h("option", {value:"training"}, "training"));
    this.mode.onchange = () => { set({mode:this.mode.value}); this.promptPanel.style.display = this.mode.value === "training" ? "none" : ""; };
    const connectBtn = h("button", {onclick: () => connect(this.url.value)}, "Connect");
    this.status = h("span", {class:"status"}, "disconnected");
    this.prompt = h("textarea", {}, "The quick brown fox"); this.ntok = inp("number", 16, {min:0}); this.temp = inp("number", 1.0, {step:0.1});
    this.topk = inp("number", 1, {min:0}); this.topp = inp("number", 0, {step:0.05, min:0, max:1});
    this.promptPanel = panel("Prompts (one per line)", this.prompt, h("div", {class:"row"}, h("div", {}, h("label", {}, "tokens"), this.ntok), h("div", {}, h("label", {}, "temperature"), this.temp)),
      h("div", {class:"row"}, h("div", {}, h("label", {}, "top-k"), this.topk), h("div", {}, h("label", {}, "top-p"), this.topp)));
    this.flags = {}; const flagRows = Object.values(FLAGS).map(n => { const c = inp("checkbox", "", {}); c.checked = !["ContextLayer_mat_mul", "MLP2_Plot"].includes(n); this.flags[n] = c; return h("div", {class:"chk"}, c, h("span", {}, FLAG_LABEL[n])); });
    this.layers = inp("text", "", {placeholder:"e.g. 1,2,8  (empty = all)"}); this.pxq = inp("number", 96); this.pxm = inp("number", 64);
    this.method = h("select", {}, ...["mean", "max", "min", "absmax", "l2"].map(m => h("option", {}, m)));
    this.pert = {}; const pertRows = ["weight", "calculation", "system"].map(k => { const on = inp("checkbox", ""), fn = h("select", {}, h("option", {value:"noise1"}, "noise1: x + c N(0,1)"), h("option", {value:"noise2"}, "noise2: x U(1-c, 1+c)")), coef = inp("number", 0.01, {step:0.01});
      this.pert[k] = {on, fn, coef}; return h("div", {}, h("div", {class:"chk"}, on, h("span", {}, k + " perturbation")), h("div", {class:"row"}, fn, coef)); });
    this.runBtn = h("button", {style:"width:100%", onclick: () => run(this.request())}, "Run");
    this.el.append(panel("Connection", h("label", {}, "WebSocket URL"), this.url, h("div", {class:"row", style:"margin-top:8px"}, this.mode, connectBtn), this.status),
      this.promptPanel, panel("Capture", ...flagRows, h("label", {}, "layers"), this.layers,
        h("div", {class:"row"}, h("div", {}, h("label", {}, "QKV pixels"), this.pxq), h("div", {}, h("label", {}, "MLP pixels"), this.pxm)), h("label", {}, "compressor"), this.method),
      panel("Perturbation config", ...pertRows), this.runBtn);
    this.watch("status");
  }
  request() {
    const layers = this.layers.value.trim() ? this.layers.value.split(",").map(s => Number(s.trim())).filter(n => n > 0) : null;
    const vis = {}; Object.entries(this.flags).forEach(([n, c]) => { vis[n] = c.checked ? "true" : "false"; if (layers) vis[n + "_layers"] = layers; });
    const dist = {}; Object.entries(this.pert).forEach(([k, p]) => { dist[k + "_perturbation"] = p.on.checked; dist[k + "_perturbation_fn"] = p.fn.value; dist[k + "_perturbation_coef"] = Number(p.coef.value); });
    const comp = {QKV:{pixels:Number(this.pxq.value), method:this.method.value}, MLP:{pixels:Number(this.pxm.value), method:this.method.value}};
    return {sampling:{prompts:this.prompt.value.split("\n").filter(s => s.length), tokens_to_generate:Number(this.ntok.value), temperature:Number(this.temp.value),
                      top_k:Number(this.topk.value), top_p:Number(this.topp.value)},
            configs:{visualization_flags:vis, disturbance_configs:dist, compressor_config:comp}};
  }
  render() { this.status.textContent = state.status; this.status.className = "status " + state.statusKind; this.runBtn.disabled = state.running; this.runBtn.textContent = state.running ? "Running..." : (state.mode === "training" ? "Run one training step" : "Generate"); }
}
