// Top-k next-token distribution for every decode step (reference components/OutputProbs.vue).
import {Component, h, panel} from "./base.js";
import {state} from "../store.js";
export class OutputProbs extends Component {
  constructor() { super(); this.tbl = h("table", {class:"probs"}); this.step = h("input", {type:"range", min:0, max:0, value:0});
    this.lab = h("small", {}); this.step.oninput = () => this.render(); this.follow = true; this.step.onchange = () => { this.follow = +this.step.value === +this.step.max; };
    this.el = panel("Output probabilities", h("div", {class:"row"}, this.step, this.lab), this.tbl); this.watch("probs"); }
  render() {
    const n = state.probs.length; this.step.max = Math.max(0, n - 1); if (this.follow) this.step.value = this.step.max;
    const i = +this.step.value; this.lab.textContent = n ? `step ${i + 1}/${n}` : "no data"; this.tbl.innerHTML = "";
    const items = state.probs[i] || []; (Array.isArray(items[0]) ? items : []).slice(0, 12).forEach(([tok, p]) => {
      const tr = this.tbl.insertRow(); tr.insertCell().textContent = String(tok); tr.insertCell().textContent = (+p).toFixed(4);
      const c = tr.insertCell(); c.className = "bar"; c.style.width = "60%"; c.appendChild(h("div", {style:`width:${Math.max(1, 100 * p)}%`})); });
  }
}
