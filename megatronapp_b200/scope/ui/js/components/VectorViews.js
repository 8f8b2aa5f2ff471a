// QKV / MLP activation views (reference QKVMatrix / QKVVector(s) / MLPVector(s).vue): token x pixel heat-map of the
// selected layer plus, for the selected token, the vector of EVERY captured layer stacked top to bottom.
import {Component, h, panel} from "./base.js";
import {state, set, tokenRows, layersWith} from "../store.js";
import {drawMatrix, legend} from "../draw.js";
import {coloredVector} from "./ColoredVector.js";

export class VectorView extends Component {
  constructor(flag, title, splitQKV = false) {
    super(); this.flag = flag; this.split = splitQKV;
    this.cv = h("canvas", {width:480, height:200}); this.leg = h("canvas", {}); this.stack = h("div", {}); this.info = h("small", {});
    this.el = panel(title, this.cv, h("div", {class:"legend"}, "−max", this.leg, "+max", this.info), this.stack);
    this.el.querySelector("h2").appendChild(h("small", {}, ""));
    legend(this.leg, false); this.watch("data", "view", "tokens");
  }
  render() {
    const rows = tokenRows(this.flag, state.layer);
    this.info.textContent = rows.length ? `  ${rows.length} tokens x ${rows[0].length} px, layer ${state.layer}` : "  no data";
    const toks = state.promptTokens.concat(state.genTokens);
    drawMatrix(this.cv, rows.length, rows.length ? rows[0].length : 0, (r, c) => rows[r][c] || 0,
               {highlightRow: state.selToken, label: (r, c) => `tok ${r}${toks[r] !== undefined ? " (" + toks[r] + ")" : ""} px ${c}`,
                onPick: r => set({selToken:r}, "tokens")});
    this.stack.innerHTML = "";
    if (state.selToken < 0) return;
    const layers = layersWith(this.flag); let mx = 1e-12; const vecs = [];
    layers.forEach(l => { const r = tokenRows(this.flag, l)[state.selToken]; if (r) { vecs.push([l, r]); r.forEach(v => mx = Math.max(mx, Math.abs(v))); } });
    vecs.forEach(([l, r]) => {
      if (this.split && r.length % 3 === 0) { const n = r.length / 3; ["Q", "K", "V"].forEach((nm, i) => this.stack.appendChild(coloredVector(r.slice(i * n, (i + 1) * n), mx, `L${l} ${nm}`))); }
      else this.stack.appendChild(coloredVector(r, mx, `layer ${l}`));
    });
  }
}
